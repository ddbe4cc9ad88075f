"""Four-index transform of a HOST-resident tensor, with the PCIe copies hidden behind the kernels.

``BasisSet(np=numpy)`` keeps ``u`` in host memory like the reference; a basis change then costs one upload and one
download of ``8 n^4`` bytes (2 x 38 ms at n = 128 over PCIe 5 x16) around 8 ms of tensor-core work.  The plain path
does upload -> four quarter GEMMs -> download back to back.  Here the same four quarter GEMMs (same contraction
order s, r, q, p as reference basis_set.py:342-348, same kernels, same rounding) are cut so that the copy engines
and the SMs work at the same time:

    upload slab a_i of u   ||  steps 1-3 on slab a_{i-1}:  u[a_i,b,c,d] -> T1[s,a_i,b,c] -> T2[r,s,a_i,b] -> T3[q,r,s,a_i]
    step 4 on column chunk p_j of C~  ||  download of chunk p_{j-1}:      T3[q,r,s,a] -> u'[p_j,q,r,s]

Steps 1-3 never mix different values of the leading index ``a`` (the sharding argument of SURVEY.md section 8e), and
step 4 -- the contraction over ``a`` -- produces the result plane by plane in ``p``, so only the last slab's three
steps and the first chunk of step 4 remain exposed.
Three CUDA streams (upload, compute = the caller's current stream, download) are ordered with events.
"""

import numpy as _numpy
import torch

from . import ops

# below this many bytes of u the plain path is used (copies too short to be worth cutting)
MIN_BYTES = 256 << 20
_SLABS = 8
_STREAMS = {}


def _side_streams(device):
    key = (device.type, device.index)
    if key not in _STREAMS:
        _STREAMS[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _STREAMS[key]


def applicable(u_host, C):
    """A C-contiguous float64 / complex128 host ndarray that is large enough.  A real tensor with an odd extent
    needs its rows padded to a 16-byte pitch before TMA can describe them; that case takes the plain path."""
    if not isinstance(u_host, _numpy.ndarray) or u_host.ndim != 4 or u_host.dtype not in (_numpy.float64, _numpy.complex128):
        return False
    if not u_host.flags.c_contiguous or u_host.nbytes < MIN_BYTES:
        return False
    return not (u_host.dtype == _numpy.float64 and u_host.shape[0] % 2)


def _chunks(total, parts):
    size = -(-total // parts)
    return [(lo, min(lo + size, total)) for lo in range(0, total, size)]


def transform_two_body(u_host, C, C_tilde=None):
    """``u'_pqrs = sum C~[p,a] C~[q,b] u[a,b,c,d] C[c,r] C[d,s]`` for a host ndarray ``u_host``; ``C`` (n, m) and
    ``C_tilde`` (m, n) or ``None`` are CUDA tensors.  Returns a pinned-memory ndarray."""
    C, C_tilde = ops._coefficients(C, C_tilde)
    n, m = C.shape
    if tuple(u_host.shape) != (n, n, n, n):
        raise ValueError(f"u must have shape {(n,) * 4} to be contracted with C {tuple(C.shape)}, got {u_host.shape}")
    device = C.device
    u_dtype = torch.complex128 if u_host.dtype == _numpy.complex128 else torch.float64
    C, C_tilde = ops.real_coefficients_if_exact(u_dtype == torch.complex128, C, C_tilde)
    t_dtype = torch.complex128 if torch.complex128 in (u_dtype, C.dtype) else torch.float64
    c_dtype = C.dtype

    compute = torch.cuda.current_stream(device)
    upload, download = _side_streams(device)
    src = torch.from_numpy(u_host)  # zero-copy view; pinned memory makes the uploads asynchronous
    result = torch.empty((m, m, m, m), dtype=t_dtype, pin_memory=True)

    slabs = _chunks(n, min(_SLABS, n))
    a_max = max(hi - lo for lo, hi in slabs)
    u_slab = [torch.empty(a_max * n**3, dtype=u_dtype, device=device) for _ in range(2)]
    t1 = torch.empty(m * a_max * n * n, dtype=t_dtype, device=device)
    t2 = torch.empty(m * m * a_max * n, dtype=t_dtype, device=device)
    t3 = torch.empty(m * m * m * n, dtype=t_dtype, device=device)

    img1 = ops.coeff_image(C, n, m, u_dtype, m, 1)
    img2 = ops.coeff_image(C, n, m, t_dtype, m, 1)
    if C_tilde is not None:
        img3 = ops.coeff_image(C_tilde, n, m, t_dtype, 1, n)
    else:
        img3 = ops.coeff_image(C, n, m, t_dtype, m, 1, conj=True)  # C~ = C^dagger, basis_set.py:338-339

    # ---- steps 1-3, slab by slab, behind the uploads ---------------------------------------------------------
    slab_free = [None, None]
    upload.wait_stream(compute)  # the slab buffers were just allocated on the compute stream
    for i, (a0, a1) in enumerate(slabs):
        A = a1 - a0
        buf = u_slab[i % 2]
        with torch.cuda.stream(upload):
            if slab_free[i % 2] is not None:
                upload.wait_event(slab_free[i % 2])
            buf[: A * n**3].view(A, n, n, n).copy_(src[a0:a1], non_blocking=True)
            arrived = torch.cuda.Event()
            arrived.record(upload)
        compute.wait_event(arrived)
        # T1[s, a_loc, b, c]: rows (a_loc, b, c), new index s slowest
        ops.quarter_transform(buf, A * n * n, n, n, img1, c_dtype, m, t1, A * n * n, 1, 0, 1, 0, A * n * n)
        slab_free[i % 2] = torch.cuda.Event()
        slab_free[i % 2].record(compute)
        # T2[r, s, a_loc, b]: rows (s, a_loc, b), new index r slowest (compact per slab)
        ops.quarter_transform(t1, m * A * n, n, n, img2, c_dtype, m, t2, m * A * n, 1, 0, 1, 0, m * A * n)
        # T3[q, r, s, a0 + a_loc]: rows (r, s, a_loc) land between the other slabs' columns of the whole T3
        ops.quarter_transform(t2, m * m * A, n, n, img3, c_dtype, m, t3[a0:], A, 1, n, 1, 0, m * m * n)

    # ---- step 4, chunk by chunk of the new leading index p, ahead of the downloads ----------------------------
    chunks = _chunks(m, min(8, max(1, m // 16)))
    p_max = max(hi - lo for lo, hi in chunks)
    out_chunk = [torch.empty(p_max * m**3, dtype=t_dtype, device=device) for _ in range(2)]
    chunk_free = [None, None]
    X = m * m * m
    for j, (p0, p1) in enumerate(chunks):
        pc = p1 - p0
        if C_tilde is not None:
            img4 = ops.coeff_image(C_tilde[p0:p1], n, pc, t_dtype, 1, n)
        else:
            img4 = ops.coeff_image(C[:, p0:p1].contiguous(), n, pc, t_dtype, pc, 1, conj=True)
        buf = out_chunk[j % 2]
        if chunk_free[j % 2] is not None:
            compute.wait_event(chunk_free[j % 2])
        ops.quarter_transform(t3, X, n, n, img4, c_dtype, pc, buf, X, 1, 0, 1, 0, X)
        done = torch.cuda.Event()
        done.record(compute)
        with torch.cuda.stream(download):
            download.wait_event(done)
            result[p0:p1].copy_(buf[: pc * X].view(pc, m, m, m), non_blocking=True)
            chunk_free[j % 2] = torch.cuda.Event()
            chunk_free[j % 2].record(download)
    download.synchronize()  # the caller receives host memory: it must be complete
    compute.wait_stream(download)
    return result.numpy()
