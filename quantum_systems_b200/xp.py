"""``xp`` -- the device array module handed to ``BasisSet(np=...)``.

The reference injects an array module as ``np`` (basis_set.py:32-38) and stores whatever that module
produces.  This module is the B200 counterpart: a numpy-*named* facade whose arrays are
``torch.Tensor`` objects resident in HBM.  It covers the names the reference's hot path and its
callers use on ``self.np`` (asarray, array, zeros, zeros_like, eye, dot, tensordot, einsum, kron,
trace, complex128, ...).  These facade functions serve the O(n^2) bookkeeping around the path and
downstream convenience; the O(n^4)/O(n^5) work never goes through them -- ``BasisSet`` routes it to
the CUDA kernels in ``libqsb200.so``.
"""

import numpy as _np
import torch as _torch

float64 = _torch.float64
complex128 = _torch.complex128
pi = _np.pi
newaxis = None

_NP_TO_TORCH = {
    _np.dtype("float64"): _torch.float64,
    _np.dtype("complex128"): _torch.complex128,
    _np.dtype("float32"): _torch.float32,
    _np.dtype("int64"): _torch.int64,
    _np.dtype("int32"): _torch.int32,
    _np.dtype("bool"): _torch.bool,
}


def device():
    """The CUDA device of this process (one process per GPU).  No CPU fallback."""
    if not _torch.cuda.is_available():
        raise RuntimeError("quantum_systems_b200.xp needs a CUDA device (there is no CPU fallback)")
    return _torch.device("cuda", _torch.cuda.current_device())


def _dtype(dtype):
    if dtype is None or isinstance(dtype, _torch.dtype):
        return dtype
    if dtype is complex:
        return _torch.complex128
    if dtype is float:
        return _torch.float64
    if dtype is int:
        return _torch.int64
    return _NP_TO_TORCH[_np.dtype(dtype)]


def is_device_array(a):
    return isinstance(a, _torch.Tensor)


def asarray(a, dtype=None):
    """Move ``a`` (tensor, ndarray, nested list, scalar) into HBM; no copy if it is already there."""
    dtype = _dtype(dtype)
    if isinstance(a, _torch.Tensor):
        t = a if a.is_cuda else a.to(device())
    elif isinstance(a, (list, tuple)) and len(a) > 0 and isinstance(a[0], _torch.Tensor):
        t = _torch.stack([asarray(x) for x in a])
    else:
        host = _np.asarray(a)
        t = _torch.from_numpy(_np.ascontiguousarray(host)).to(device())
    return t if dtype is None or t.dtype == dtype else t.to(dtype)


array = asarray


def asnumpy(a):
    """Device tensor -> host ndarray (synchronising copy)."""
    if isinstance(a, _torch.Tensor):
        return a.detach().cpu().numpy()
    return _np.asarray(a)


def zeros(shape, dtype=float64):
    return _torch.zeros(shape, dtype=_dtype(dtype), device=device())


def ones(shape, dtype=float64):
    return _torch.ones(shape, dtype=_dtype(dtype), device=device())


def empty(shape, dtype=float64):
    return _torch.empty(shape, dtype=_dtype(dtype), device=device())


def zeros_like(a, dtype=None):
    return _torch.zeros_like(asarray(a), dtype=_dtype(dtype))


def eye(n, dtype=float64):
    return _torch.eye(n, dtype=_dtype(dtype), device=device())


def diag(a):
    return _torch.diag(asarray(a))


def arange(*args, dtype=None):
    return _torch.arange(*args, dtype=_dtype(dtype), device=device())


def linspace(start, stop, num, dtype=float64):
    return _torch.linspace(start, stop, num, dtype=_dtype(dtype), device=device())


def conj(a):
    return _torch.conj(asarray(a)).resolve_conj()


conjugate = conj


def real(a):
    return _torch.real(asarray(a))


def imag(a):
    a = asarray(a)
    return _torch.imag(a) if a.is_complex() else _torch.zeros_like(a)


def abs(a):  # noqa: A001 - numpy name
    return _torch.abs(asarray(a))


def sqrt(a):
    return _torch.sqrt(asarray(a))


def exp(a):
    return _torch.exp(asarray(a))


def sum(a, axis=None):  # noqa: A001 - numpy name
    a = asarray(a)
    return a.sum() if axis is None else a.sum(dim=axis)


def transpose(a, axes=None):
    a = asarray(a)
    if axes is None:
        axes = tuple(reversed(range(a.dim())))
    return a.permute(*axes)


def swapaxes(a, i, j):
    return _torch.swapaxes(asarray(a), i, j)


def moveaxis(a, src, dst):
    return _torch.movedim(asarray(a), src, dst)


def _promote(a, b):
    a, b = asarray(a), asarray(b)
    dt = _torch.promote_types(a.dtype, b.dtype)
    return a.to(dt), b.to(dt)


def dot(a, b):
    a, b = _promote(a, b)
    return _torch.matmul(a, b)


matmul = dot


def tensordot(a, b, axes=2):
    a, b = _promote(a, b)
    if isinstance(axes, int):
        return _torch.tensordot(a, b, dims=axes)
    ax_a, ax_b = axes
    ax_a = [ax_a] if isinstance(ax_a, int) else list(ax_a)
    ax_b = [ax_b] if isinstance(ax_b, int) else list(ax_b)
    return _torch.tensordot(a, b, dims=(ax_a, ax_b))


def einsum(subscripts, *operands, **_ignored):
    ops_ = [asarray(o) for o in operands]
    dt = ops_[0].dtype
    for o in ops_[1:]:
        dt = _torch.promote_types(dt, o.dtype)
    return _torch.einsum(subscripts.replace(" ", ""), *[o.to(dt) for o in ops_])


def kron(a, b):
    a, b = _promote(a, b)
    return _torch.kron(a.contiguous(), b.contiguous())


def trace(a, offset=0, axis1=0, axis2=1):
    return _torch.diagonal(asarray(a), offset=offset, dim1=axis1, dim2=axis2).sum(-1)


def allclose(a, b, rtol=1e-5, atol=1e-8):
    a, b = _promote(a, b)
    return bool(_torch.allclose(a, b, rtol=rtol, atol=atol))


def copy(a):
    return asarray(a).clone()


def astype(a, dtype):
    return asarray(a).to(_dtype(dtype))


class _Linalg:
    @staticmethod
    def inv(a):
        return _torch.linalg.inv(asarray(a))

    @staticmethod
    def eigh(a):
        w, v = _torch.linalg.eigh(asarray(a))
        return w, v

    @staticmethod
    def qr(a):
        q, r = _torch.linalg.qr(asarray(a))
        return q, r

    @staticmethod
    def norm(a):
        return _torch.linalg.norm(asarray(a))


linalg = _Linalg()


class _Random:
    """Host-seeded random numbers (numpy's global stream, like the reference's RandomBasisSet,
    random_basis.py:53-69), then moved to the device."""

    @staticmethod
    def random(shape=None):
        r = _np.random.random(shape)
        return asarray(r) if shape is not None else r

    @staticmethod
    def choice(a):
        return _np.random.choice(a)

    @staticmethod
    def seed(s):
        _np.random.seed(s)


random = _Random()
