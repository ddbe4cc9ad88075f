"""GPU parity of the sharded schedule.  On one GPU all W ranks are driven by one process
(``EmulatedContext``): the scattering quarter GEMM writes into W destination buffers exactly as it
writes into W peers' IPC mappings.  With >= 2 GPUs the real one-process-per-GPU path (CUDA IPC peer
stores over NVLink, NCCL barriers) is run under torchrun by tests/_multi_gpu_worker.py."""

import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import assert_close_scaled
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if complex_ else x


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n,m,u_complex,c_complex,biorth", [
    (16, 16, False, False, False),
    (9, 9, False, False, False),
    (13, 20, False, False, False),
    (20, 12, True, True, True),
    (12, 12, False, True, False),
    (40, 40, False, False, False),
    (24, 24, True, False, False),   # complex u, real C: the split (2M) quarter GEMM, scattering and local
    (14, 18, True, False, True),
])
def test_emulated_sharded_transform(world, n, m, u_complex, c_complex, biorth):
    from quantum_systems_b200 import sharded

    rng = np.random.default_rng(100 * n + m + world)
    u = rand(rng, (n,) * 4, u_complex)
    C = rand(rng, (n, m), c_complex)
    Ct = rand(rng, (m, n), c_complex) if biorth else None
    ctx = sharded.EmulatedContext(world)
    basis = sharded.ShardedBasisSet.from_global(ctx, np.eye(n), np.eye(n), u)
    out = sharded.transform_two_body_sharded(basis.u, dev(C), None if Ct is None else dev(Ct))
    expected = oracle.transform_two_body_elements(u, C, Ct)
    got = out.gather().cpu().numpy()
    assert got.dtype == expected.dtype
    assert_close_scaled(got, expected, rel=1e-12)
    if m == n and out.dtype == basis.u.dtype:
        out2 = sharded.transform_two_body_sharded(out, dev(C), None if Ct is None else dev(Ct))
        assert_close_scaled(out2.gather().cpu().numpy(), oracle.transform_two_body_elements(expected, C, Ct), rel=1e-11)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("out_complex", [False, True])
def test_sharded_spin_doubling_change_basis_fock(world, out_complex):
    """add_spin + anti_symmetrize on plane ranges, sharded change_basis, row-sharded Fock matrix."""
    from quantum_systems_b200 import sharded

    rng = np.random.default_rng(world)
    l, n_occ = 7, 4
    u = rng.standard_normal((l,) * 4)
    h = rng.standard_normal((l, l))
    s = np.eye(l)
    ctx = sharded.EmulatedContext(world)
    basis = sharded.ShardedBasisSet.from_spatial(
        ctx, h, s, u, anti_symmetrize=True, out_dtype=torch.complex128 if out_complex else torch.float64
    )
    ref_u = oracle.anti_symmetrize_u(oracle.add_spin_two_body(u))
    np.testing.assert_array_equal(basis.u.gather().cpu().numpy(), ref_u.astype(np.complex128 if out_complex else np.float64))
    ref_h = oracle.add_spin_one_body(h)
    C = np.linalg.qr(rng.standard_normal((2 * l, 2 * l)))[0]
    basis.change_basis(dev(C))
    ref_u2 = oracle.transform_two_body_elements(ref_u, C)
    ref_h2 = oracle.transform_one_body_elements(ref_h, C)
    assert_close_scaled(basis.u.gather().cpu().numpy(), ref_u2, rel=1e-12)
    assert_close_scaled(basis.h.cpu().numpy(), ref_h2, rel=1e-12)
    f = basis.construct_fock_matrix(basis.h, basis.u, n_occ)
    assert_close_scaled(f.cpu().numpy(), oracle.construct_fock_matrix_general(ref_h2, ref_u2, n_occ), rel=1e-12)


def test_scatter_rejects_too_many_destinations():
    from quantum_systems_b200 import sharded

    with pytest.raises(RuntimeError, match="at most"):
        ctx = sharded.EmulatedContext(17)
        basis = sharded.ShardedBasisSet.from_global(ctx, np.eye(34), np.eye(34), np.zeros((34,) * 4))
        sharded.transform_two_body_sharded(basis.u, dev(np.eye(34)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["peer", "collective"])
def test_two_gpu_processes(exchange):
    """One process per GPU under torchrun: CUDA IPC peer stores (or NCCL all_to_all) + NCCL barriers."""
    world = min(torch.cuda.device_count(), 8)
    world = 1 << (world.bit_length() - 1)
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "_multi_gpu_worker.py"),
        exchange,
    ]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "MULTI_GPU_OK" in proc.stdout


@pytest.mark.parametrize("l,G", [(6, 101), (7, 64), (10, 201)])
def test_odqd_build_by_planes(l, G):
    """The grid Coulomb build restricted to leading-index planes tiles the full build."""
    from quantum_systems_b200 import ops

    grid, eps, C = oracle.odqd_orbitals(l, 5.0, G, lambda x: 0.5 * x**2)
    expected = np.ascontiguousarray(oracle.odqd_coulomb_elements(C, grid, 1.0, 0.25))
    C_dev, grid_dev = dev(C), dev(grid[1:-1])
    whole = ops.odqd_coulomb(C_dev, grid_dev, 1.0, 0.25)
    assert_close_scaled(whole.cpu().numpy(), expected)
    cuts = [0, 1, 1, l // 2, l]
    parts = [ops.odqd_coulomb(C_dev, grid_dev, 1.0, 0.25, planes=(a0, a1)) for a0, a1 in zip(cuts, cuts[1:])]
    assert parts[1].shape[0] == 0
    assert_close_scaled(torch.cat(parts).cpu().numpy(), expected)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("out_complex", [False, True])
def test_sharded_odqd_pipeline_without_a_replicated_tensor(world, out_complex):
    """ODQD build -> add_spin + anti-symmetrise -> change_basis, every stage sharded (emulated ranks): no rank holds
    the spatial l^4 tensor, each builds only the planes behind its own spin-orbital planes."""
    from quantum_systems_b200 import sharded

    l, G = 7, 101
    grid, eps, C = oracle.odqd_orbitals(l, 5.0, G, lambda x: 0.5 * x**2)
    u_spatial = oracle.odqd_coulomb_elements(C, grid, 1.0, 0.25)
    h = np.diag(eps)
    asked = []

    def planes(p0, p1, build=sharded.odqd_spatial_planes(C, grid[1:-1], 1.0, 0.25)):
        asked.append((p0, p1))
        return build(p0, p1)

    ctx = sharded.EmulatedContext(world)
    dtype = torch.complex128 if out_complex else torch.float64
    basis = sharded.ShardedBasisSet.from_spatial_planes(ctx, h, np.eye(l), l, planes, out_dtype=dtype)
    expected = oracle.anti_symmetrize_u(oracle.add_spin_two_body(u_spatial))
    got = basis.u.gather().cpu().numpy()
    assert got.dtype == (np.complex128 if out_complex else np.float64)
    assert_close_scaled(got, expected.astype(got.dtype))
    assert all(p1 - p0 <= -(-(2 * l) // world) // 2 + 1 for p0, p1 in asked)  # only the needed planes were built
    rng = np.random.default_rng(world)
    Cb = np.linalg.qr(rng.standard_normal((2 * l, 2 * l)))[0]
    basis.change_basis(dev(Cb))
    assert_close_scaled(basis.u.gather().cpu().numpy(), oracle.transform_two_body_elements(expected, Cb).astype(got.dtype))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n,m,u_complex,c_complex,biorth", [
    (48, 48, False, False, False),
    (50, 56, False, False, False),
    (56, 48, True, True, True),
    (48, 48, True, False, False),   # split (2M) kernel under the symmetric schedule
    (49, 49, False, False, False),  # odd real extent: padded pitch in the packed pair layout
])
def test_emulated_symmetry_aware_sharded_transform(world, n, m, u_complex, c_complex, biorth):
    """Exactly anti-symmetric u on an (emulated) sharded tensor: detection on every rank's slab, steps 3-4 on the
    pairs of the cyclic rule packed by pair, scattering store through the pair table, local mirror fill."""
    from quantum_systems_b200 import sharded

    rng = np.random.default_rng(100 * n + m + world)
    u = rand(rng, (n,) * 4, u_complex)
    u = u - u.transpose(0, 1, 3, 2)
    C = rand(rng, (n, m), c_complex)
    Ct = rand(rng, (m, n), c_complex) if biorth else None
    ctx = sharded.EmulatedContext(world)
    basis = sharded.ShardedBasisSet.from_global(ctx, np.eye(n), np.eye(n), u)
    assert sharded.is_antisymmetric_last_pair(basis.u)
    expected = oracle.transform_two_body_elements(u, C, Ct)
    out = sharded.transform_two_body_sharded(basis.u, dev(C), None if Ct is None else dev(Ct))
    got = out.gather().cpu().numpy()
    assert_close_scaled(got, expected, rel=1e-12)
    np.testing.assert_array_equal(got, -got.transpose(0, 1, 3, 2))  # the mirror image is a copy
    plain = sharded.transform_two_body_sharded(basis.u, dev(C), None if Ct is None else dev(Ct), symmetry=0)
    assert_close_scaled(plain.gather().cpu().numpy(), expected, rel=1e-12)


@pytest.mark.parametrize("world,complex_,l", [(4, False, 26), (3, True, 26), (2, False, 66), (8, False, 68)])
def test_sharded_odqd_general_system_against_the_closed_form(world, complex_, l):
    """The pipeline of BASELINE configs[3]/[4] in miniature (emulated ranks): ShardedBasisSet.from_odqd (sharded grid
    build -> add_spin + anti-symmetrise) -> change_basis twice -> Fock matrix, checked against the closed form of
    the ODQD-structured tensor (tests/closed_form.py) exactly as bench.py checks the full-size runs, and against
    the oracle's explicit pipeline."""
    import closed_form
    from quantum_systems_b200 import sharded
    from quantum_systems_b200.potentials import HOPotential

    G = 4 * l + 1  # l = 66, 68: several column tiles per launch, so the first exchange really skips tiles
    n = 2 * l
    dtype = torch.complex128 if complex_ else torch.float64
    ctx = sharded.EmulatedContext(world)
    basis = sharded.ShardedBasisSet.from_odqd(ctx, l, 8.0, G, potential=HOPotential(0.5), out_dtype=dtype)
    assert basis.u.proven_antisymmetric and basis.includes_spin and basis.l == n
    rng = np.random.default_rng(world)
    if complex_:
        C = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + 4 * np.eye(n)
        Ct = np.linalg.inv(C)
    else:
        C, Ct = np.linalg.qr(rng.standard_normal((n, n)))[0], None
    W = closed_form.shielded_coulomb(basis.inner_grid, 1.0, 0.25)
    # explicit reference pipeline at this size
    grid, eps, Lg = oracle.odqd_orbitals(l, 8.0, G, HOPotential(0.5))
    np.testing.assert_allclose(np.abs(Lg), np.abs(basis.grid_coefficients), atol=1e-12)
    small = l <= 30  # the explicit oracle pipeline next to the closed form (seconds at n = 52, minutes at n = 132)
    if small:
        spin = oracle.anti_symmetrize_u(oracle.add_spin_two_body(np.ascontiguousarray(
            oracle.odqd_coulomb_elements(basis.grid_coefficients, grid, 1.0, 0.25))))
        assert_close_scaled(basis.u.gather().cpu().numpy(), spin.astype(np.complex128 if complex_ else np.float64))

    basis.change_basis(dev(C), None if Ct is None else dev(Ct))
    form = closed_form.SpinDoubledClosedForm(basis.grid_coefficients, W, C, Ct)
    for r in range(world):
        p0, p1 = basis.u.planes(r)
        errs = closed_form.check_shard(closed_form.TorchSlab(basis.u.local(r)), p0, form, np.random.default_rng(r), samples=500)
        assert max(errs[:2]) <= 1e-12 * errs[3] and errs[2] == 0.0
    if small:
        assert_close_scaled(basis.u.gather().cpu().numpy(), oracle.transform_two_body_elements(spin, C, Ct))
    # second call of the chain: no new symmetry test, net transform C C / C~ C~
    basis.change_basis(dev(C), None if Ct is None else dev(Ct))
    Ct1 = C.conj().T if Ct is None else Ct
    net = closed_form.SpinDoubledClosedForm(basis.grid_coefficients, W, C @ C, Ct1 @ Ct1)
    errs = closed_form.check_shard(closed_form.TorchSlab(basis.u.local(0)), 0, net, np.random.default_rng(7), samples=500)
    assert max(errs[:2]) <= 1e-11 * errs[3]
    n_occ = 6
    f = basis.construct_fock_matrix(basis.h, basis.u, n_occ).cpu().numpy()
    ref_f = basis.h.cpu().numpy() + net.fock_two_body(n_occ)
    assert_close_scaled(f, ref_f, rel=1e-11)
