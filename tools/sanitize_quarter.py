"""Small-size driver for `compute-sanitizer` (memcheck / racecheck / synccheck) over every variant of the quarter
GEMM: all column-tile counts NT = 1..8 (real) and the complex 4M image, the split (2M) kernel NTC = 1..4, masked
launches with tile lists and packed pair tables (symmetry-aware transform, both symmetries), the scattering store
with column dealing through an emulated 2- and 3-rank context (peer stores to local memory), the row-table
scattering store, and ragged extents (odd n, short last K chunk, partial row tiles).  Every result is compared with
a plain torch einsum so that a sanitizer run is also a correctness run.

    compute-sanitizer --tool memcheck  python tools/sanitize_quarter.py
    compute-sanitizer --tool racecheck python tools/sanitize_quarter.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quantum_systems_b200 import _native, ops, sharded


def reference(u, C, Ct=None):
    Ct = C.conj().T if Ct is None else Ct
    dt = torch.promote_types(u.dtype, C.dtype)
    return torch.einsum("pa,qb,abcd,cr,ds->pqrs", Ct.to(dt), Ct.to(dt), u.to(dt), C.to(dt), C.to(dt))


def close(got, ref, what):
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, f"{what}: {err:.2e}"


def main():
    torch.manual_seed(0)
    lib = _native.load()
    n0 = lib.qs_launch_count()
    f64, c128 = torch.float64, torch.complex128
    rnd = lambda shape, dt: torch.randn(shape, dtype=dt, device="cuda")  # noqa: E731

    # generic kernel: W' = 8 NT real columns for NT = 1..8, ragged rows, short last chunk (K = 12, 20), odd extents
    for n, m in [(12, 8), (12, 16), (13, 24), (20, 32), (9, 40), (12, 48), (12, 56), (12, 64), (11, 70)]:
        u, C = rnd((n,) * 4, f64), rnd((n, m), f64)
        close(ops.transform_two_body(u, C, symmetry=0), reference(u, C), f"real {n}->{m}")
    # complex 4M image (real u x complex C, complex u x complex C, bi-orthogonal) and the split 2M kernel NTC = 1..4
    for n, m in [(10, 4), (10, 12), (9, 20), (12, 30)]:
        uc, Cc, Ct = rnd((n,) * 4, c128), rnd((n, m), c128), rnd((m, n), c128)
        close(ops.transform_two_body(uc, Cc, Ct, symmetry=0), reference(uc, Cc, Ct), f"complex {n}->{m}")
        close(ops.transform_two_body(uc.real.contiguous(), Cc, symmetry=0), reference(uc.real, Cc), f"real x complex {n}->{m}")
        Cr = rnd((n, m), f64)
        close(ops.transform_two_body(uc, Cr, symmetry=0), reference(uc, Cr), f"split {n}->{m}")
    # masked launches: tile lists, packed pair layout, mirror fill (both symmetries, real and complex, rectangular)
    for n, m, dt in [(48, 48, f64), (50, 56, f64), (48, 50, c128)]:
        u = rnd((n,) * 4, dt)
        C = rnd((n, m), dt)
        anti = u - u.transpose(2, 3)
        exch = 0.5 * (u + u.permute(1, 0, 3, 2))
        close(ops.transform_two_body(anti.contiguous(), C, symmetry=1), reference(anti, C), f"antisym {n}->{m}")
        close(ops.transform_two_body(exch.contiguous(), C, symmetry=2), reference(exch, C), f"exchange {n}->{m}")
    # scattering store (fused re-partition) with column dealing, emulated ranks; symmetric variant with row tables
    for world, n, dt in [(2, 24, f64), (3, 20, c128), (2, 48, f64)]:
        ctx = sharded.EmulatedContext(world)
        u = rnd((n,) * 4, dt)
        u = (u - u.transpose(2, 3)).contiguous()
        C = rnd((n, n), dt)
        basis = sharded.ShardedBasisSet.from_global(ctx, torch.eye(n, dtype=f64), torch.eye(n, dtype=f64), u)
        for symmetry in (0, None):
            out = sharded.transform_two_body_sharded(basis.u, C, symmetry=symmetry)
            close(out.gather(), reference(u, C), f"sharded world={world} n={n} symmetry={symmetry}")
    # one-body and grid-function transforms (X = n rows: a single partial tile)
    h, C = rnd((17, 17), c128), rnd((17, 9), c128)
    close(ops.transform_one_body(h, C), C.conj().T @ h @ C, "one-body")
    torch.cuda.synchronize()
    print("SANITIZE_DRIVER_OK launches", lib.qs_launch_count() - n0)


if __name__ == "__main__":
    main()
