"""Array-module plumbing behind the ``np`` hook.

Two storage modes, selected by the module injected as ``np`` (reference basis_set.py:32-38, :272-296):

* ``quantum_systems_b200.xp`` (default): arrays are torch tensors resident in HBM; nothing crosses
  PCIe between hot-path calls.
* ``numpy``: arrays are host ndarrays exactly as in the reference; every hot-path method stages its
  operands into HBM, runs the CUDA kernels and copies the result back (pinned staging both ways).

Either way all arithmetic of the hot path runs in ``libqsb200.so`` on the GPU.
"""

import numpy as _numpy
import torch

from . import xp as _xp


def is_host_module(module):
    """True when ``module`` is numpy (host storage); anything else is treated as device storage."""
    return module is _numpy or getattr(module, "__name__", "") == "numpy"


def default_module():
    return _xp


def to_device(a, dtype=None):
    """Any array-like -> contiguous float64/complex128 CUDA tensor (H2D copy for host data)."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(_xp.device(), non_blocking=True)
    else:
        host = _numpy.asarray(a)
        if host.dtype.kind in "iub":
            host = host.astype(_numpy.float64)
        elif host.dtype == _numpy.float32:
            host = host.astype(_numpy.float64)
        elif host.dtype == _numpy.complex64:
            host = host.astype(_numpy.complex128)
        if not host.flags.c_contiguous:
            host = _numpy.ascontiguousarray(host)
        if not host.flags.writeable:
            host = host.copy()
        t = torch.from_numpy(host).to(_xp.device(), non_blocking=True)
    if t.dtype not in (torch.float64, torch.complex128):
        t = t.to(torch.complex128 if t.is_complex() else torch.float64)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


def to_host(t):
    """CUDA tensor -> ndarray through a pinned staging buffer (the ndarray keeps the buffer alive)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        return _numpy.asarray(t)
    if not t.is_cuda:
        return t.numpy()
    staging = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    staging.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return staging.numpy()


def to_module(t, module):
    """Hand a result to the caller in the storage type of ``module``."""
    if t is None:
        return None
    if is_host_module(module):
        return to_host(t)
    return t


def store(a, module):
    """Convert an incoming array to the storage type of ``module`` (used by setters/change_module)."""
    if a is None:
        return None
    if is_host_module(module):
        return to_host(a) if isinstance(a, torch.Tensor) else _numpy.asarray(a)
    return to_device(a)
