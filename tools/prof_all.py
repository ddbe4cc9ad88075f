"""One launch of every kernel family of libqsb200.so at a representative size -- the driver of the per-kernel
`ncu --set full` capture summarised under profiles/ (tools/ncu_summary.py).  Prints the launch count."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quantum_systems_b200 import _native, ops, two_dim_ho


def main():
    torch.manual_seed(0)
    lib = _native.load()
    n0 = lib.qs_launch_count()
    f64, c128 = torch.float64, torch.complex128

    # four-index transform: real (generic kernel), complex x complex (4M image), complex u x real C (split, 2M)
    n = 96
    u = torch.randn((n,) * 4, dtype=f64, device="cuda")
    C = torch.linalg.qr(torch.randn((n, n), dtype=f64, device="cuda"))[0].contiguous()
    ops.transform_two_body(u, C)
    n = 64
    uc = torch.randn((n,) * 4, dtype=c128, device="cuda")
    Cc = torch.linalg.qr(torch.randn((n, n), dtype=c128, device="cuda"))[0].contiguous()
    ops.transform_two_body(uc, Cc)
    ops.transform_two_body(uc, Cc.real.contiguous())
    h = torch.randn((n, n), dtype=c128, device="cuda")
    ops.transform_one_body(h, Cc)

    # spin doubling fused with anti-symmetrisation (+ cast), stand-alone anti-symmetrisation, one-body kron
    l = 48
    us = torch.randn((l,) * 4, dtype=f64, device="cuda")
    a = ops.add_spin_two_body(us, anti_symmetrize=True, out_dtype=f64)
    ops.add_spin_two_body(us, anti_symmetrize=True, out_dtype=c128)
    ops.anti_symmetrize(a)
    ops.add_spin_one_body(torch.randn((l, l), dtype=f64, device="cuda"), out_dtype=c128)
    sx = torch.randn((3, 48, 48), dtype=c128, device="cuda")
    ops.spin_squared_two_body(sx[0], sx[1], sx[2], anti_symmetrize=True)

    # Fock matrices, reference-energy traces, o/v block extraction, scaled sums
    n, n_occ = 96, 10
    hr = torch.randn((n, n), dtype=f64, device="cuda")
    ops.fock_general(hr, a, n_occ)
    ops.fock_spatial(hr, a, n_occ)
    ops.occupied_traces(hr, a, n_occ)
    ops.extract_block(a, slice(0, n_occ), slice(0, n_occ), slice(n_occ, n), slice(n_occ, n))
    ops.extract_block(a, slice(n_occ, n), slice(n_occ, n), slice(n_occ, n), slice(n_occ, n))
    ops.scale_add(a, 0.5)
    ops.scale_add(a, 0.5, a, 2.0)

    # grid builders: ODQD shielded Coulomb (config 3 shape, reduced), 2-D oscillator elements, sinc-DVR transform
    G, lq = 1001, 48
    Cg = torch.randn((G - 2, lq), dtype=f64, device="cuda")
    grid = torch.linspace(-10, 10, G, dtype=f64, device="cuda")[1:-1].contiguous()
    ops.odqd_coulomb(Cg, grid, 1.0, 0.25)
    for lt in (36, 66):
        nm = np.array([two_dim_ho.get_indices_nm(p) for p in range(lt)], dtype=np.int64)
        ops.tdho_coulomb(nm[:, 0], nm[:, 1])
    w = torch.randn((64, 64), dtype=f64, device="cuda")
    ops.transform_two_body_diagonal(w, torch.randn((64, 48), dtype=c128, device="cuda"), anti_symmetrize=True)
    torch.cuda.synchronize()
    print("launches", lib.qs_launch_count() - n0)


if __name__ == "__main__":
    main()
