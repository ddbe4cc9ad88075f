"""torchrun worker of tests/test_gpu_sharded.py::test_two_gpu_processes (one rank per GPU)."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import qs_oracle as oracle  # noqa: E402  (checker)
from quantum_systems_b200 import sharded  # noqa: E402


def main():
    exchange = sys.argv[1] if len(sys.argv) > 1 else "peer"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ctx = sharded.ProcessContext(exchange=exchange)
    try:
        for n, m, complex_ in [(24, 24, False), (18, 26, True), (27, 27, False)]:
            rng = np.random.default_rng(n)
            u = rng.standard_normal((n,) * 4) + (1j * rng.standard_normal((n,) * 4) if complex_ else 0)
            C = rng.standard_normal((n, m)) + (1j * rng.standard_normal((n, m)) if complex_ else 0)
            basis = sharded.ShardedBasisSet.from_global(ctx, np.eye(n), np.eye(n), u)
            C_dev = torch.from_numpy(C).cuda()
            out = sharded.transform_two_body_sharded(basis.u, C_dev)
            expected = oracle.transform_two_body_elements(u, C)
            q0, q1 = out.planes(rank)
            got = out.local().cpu().numpy()
            err = np.abs(got - expected[q0:q1]).max() if q1 > q0 else 0.0
            assert err <= 1e-12 * np.abs(expected).max(), f"rank {rank} n={n}: {err}"
            if m == n:
                out2 = sharded.transform_two_body_sharded(out, C_dev)
                expected2 = oracle.transform_two_body_elements(expected, C)
                full = out2.gather().cpu().numpy()
                assert np.abs(full - expected2).max() <= 1e-11 * np.abs(expected2).max()
        # anti-symmetric input above the detection threshold: the symmetry-aware schedule over real peer memory
        n = 48
        rng = np.random.default_rng(48)
        u = rng.standard_normal((n,) * 4)
        u = u - u.transpose(0, 1, 3, 2)
        C = np.linalg.qr(rng.standard_normal((n, n)))[0]
        basis = sharded.ShardedBasisSet.from_global(ctx, np.eye(n), np.eye(n), u)
        assert sharded.is_antisymmetric_last_pair(basis.u)
        out = sharded.transform_two_body_sharded(basis.u, torch.from_numpy(C).cuda())
        assert out.proven_antisymmetric == (exchange == "peer")
        expected = oracle.transform_two_body_elements(u, C)
        full = out.gather().cpu().numpy()
        assert np.abs(full - expected).max() <= 1e-12 * np.abs(expected).max()
        if exchange == "peer":  # the symmetry-aware schedule writes the mirror image: exact negatives
            assert np.array_equal(full, -full.transpose(0, 1, 3, 2))
        else:  # the collective schedule runs the four full steps: anti-symmetric to rounding
            assert np.abs(full + full.transpose(0, 1, 3, 2)).max() <= 1e-12 * np.abs(expected).max()
        # spin doubling + change_basis + Fock through the container
        rng = np.random.default_rng(5)
        l, n_occ = 8, 4
        u = rng.standard_normal((l,) * 4)
        h = rng.standard_normal((l, l))
        basis = sharded.ShardedBasisSet.from_spatial(ctx, h, np.eye(l), u, out_dtype=torch.float64)
        C = np.linalg.qr(rng.standard_normal((2 * l, 2 * l)))[0]
        basis.change_basis(torch.from_numpy(C).cuda())
        ref_u = oracle.transform_two_body_elements(oracle.anti_symmetrize_u(oracle.add_spin_two_body(u)), C)
        ref_h = oracle.transform_one_body_elements(oracle.add_spin_one_body(h), C)
        f = basis.construct_fock_matrix(basis.h, basis.u, n_occ).cpu().numpy()
        ref_f = oracle.construct_fock_matrix_general(ref_h, ref_u, n_occ)
        assert np.abs(f - ref_f).max() <= 1e-12 * np.abs(ref_f).max()
        # consumers of the sharded result: replicated o/v blocks, reference energy, scaled copy
        o, v = slice(0, n_occ), slice(n_occ, 2 * l)
        scale = np.abs(ref_u).max()
        assert np.abs(basis.u.extract(o, o, v, v).cpu().numpy() - ref_u[o, o, v, v]).max() <= 1e-12 * scale
        assert np.abs(basis.u.extract(v, o, None, o).cpu().numpy() - ref_u[v, o, :, o]).max() <= 1e-12 * scale
        e_ref = oracle.reference_energy_general(ref_h, ref_u, n_occ)
        assert abs(basis.compute_reference_energy(n_occ) - e_ref) <= 1e-11 * abs(e_ref)
        q0, q1 = basis.u.planes(rank)
        doubled = basis.u.copy().axpby_(2.0)
        assert np.abs(doubled.local().cpu().numpy() - 2.0 * ref_u[q0:q1]).max() <= 2e-12 * scale
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print("MULTI_GPU_OK", exchange, "world", world, flush=True)
    finally:
        if exchange == "peer":
            ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
