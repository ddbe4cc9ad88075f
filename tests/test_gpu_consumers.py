"""GPU parity of the consumer passes: u[o,o,v,v]-style block extraction, the
``alpha x + beta y`` pass behind h_t / u_t, and the occupied traces behind compute_reference_energy -- on one
GPU and on an (emulated) sharded tensor -- against numpy slicing and the oracle.  Data movement is held to
bit equality, sums to 1e-13 relative."""

import numpy as np
import pytest

from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if complex_ else x


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("n,n_occ", [(6, 2), (13, 5), (32, 10)])
def test_extract_block_matches_numpy_slicing(n, n_occ, complex_):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(n)
    u = rand(rng, (n,) * 4, complex_)
    u_dev = dev(u)
    o, v = slice(0, n_occ), slice(n_occ, n)
    for sl in [(o, o, v, v), (v, v, o, o), (o, v, o, v), (o, o, o, o), (v, v, v, v), (None, o, None, o),
               (slice(1, 4), slice(2, 2), o, v), (slice(-3, None), slice(None, -1), None, slice(1, 2))]:
        expected = u[tuple(slice(None) if s is None else s for s in sl)]
        got = ops.extract_block(u_dev, *sl)
        assert tuple(got.shape) == expected.shape
        np.testing.assert_array_equal(got.cpu().numpy(), expected)
    with pytest.raises(ValueError):
        ops.extract_block(u_dev, slice(0, n, 2))


@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("count", [1, 2, 7, 1000, 65537])
def test_scale_add(count, complex_):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(count)
    x, y = rand(rng, (count,), complex_), rand(rng, (count,), complex_)
    np.testing.assert_array_equal(ops.scale_add(dev(x), 0.75).cpu().numpy(), 0.75 * x)
    got = ops.scale_add(dev(x), 2.0, dev(y), -0.5).cpu().numpy()
    np.testing.assert_allclose(got, 2.0 * x - 0.5 * y, rtol=1e-15, atol=1e-15)
    got = ops.scale_add(dev(x), 1 + 2j, dev(y), -0.5j).cpu().numpy()  # complex factors widen a real tensor
    assert got.dtype == np.complex128
    np.testing.assert_allclose(got, (1 + 2j) * x - 0.5j * y, rtol=1e-15, atol=1e-15)
    x_dev = dev(x)
    assert ops.scale_add(x_dev, 3.0, out=x_dev) is x_dev  # in place
    np.testing.assert_array_equal(x_dev.cpu().numpy(), 3.0 * x)


@pytest.mark.parametrize("h_complex,u_complex", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("n,n_occ", [(5, 0), (6, 3), (20, 20), (40, 11)])
def test_occupied_traces_and_reference_energies(n, n_occ, h_complex, u_complex):
    from quantum_systems_b200 import BasisSet, GeneralOrbitalSystem, SpatialOrbitalSystem, ops, xp

    rng = np.random.default_rng(n + n_occ)
    h, u = rand(rng, (n, n), h_complex), rand(rng, (n,) * 4, u_complex)
    if n_occ:
        got = ops.occupied_traces(dev(h), dev(u), n_occ).cpu().numpy()
        o = slice(0, n_occ)
        expected = [np.trace(h[o, o]), np.einsum("ijij->", u[o, o, o, o]), np.einsum("ijji->", u[o, o, o, o])]
        np.testing.assert_allclose(got, expected, rtol=1e-13, atol=1e-13)

    for module in (np, xp):
        bs = BasisSet(n, 1, np=module, includes_spin=True, anti_symmetrized_u=True)
        bs.h, bs.u, bs.s = h.copy(), u.copy(), np.eye(n)
        bs.nuclear_repulsion_energy = 0.5
        gos = GeneralOrbitalSystem(n_occ, bs)
        np.testing.assert_allclose(
            gos.compute_reference_energy(), oracle.reference_energy_general(h, u, n_occ, 0.5), rtol=1e-13
        )
        if n_occ % 2 == 0 and 2 * (n_occ // 2) <= n:
            bs2 = BasisSet(n, 1, np=module)
            bs2.h, bs2.u, bs2.s = h.copy(), u.copy(), np.eye(n)
            spas = SpatialOrbitalSystem(n_occ, bs2)
            np.testing.assert_allclose(
                spas.compute_reference_energy(), oracle.reference_energy_spatial(h, u, n_occ // 2), rtol=1e-13
            )
        if not (h_complex or u_complex):
            assert isinstance(gos.compute_reference_energy(), float)  # real integrals give a real energy


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("complex_", [False, True])
def test_consumers_of_an_emulated_sharded_tensor(world, complex_):
    from quantum_systems_b200 import sharded

    n, n_occ = 11, 4
    rng = np.random.default_rng(31 + world)
    u, h = rand(rng, (n,) * 4, complex_), rand(rng, (n, n), complex_)
    ctx = sharded.EmulatedContext(world)
    basis = sharded.ShardedBasisSet.from_global(ctx, h, np.eye(n), u, includes_spin=True)
    o, v = slice(0, n_occ), slice(n_occ, n)
    for sl in [(o, o, v, v), (v, o, v, o), (o, o, o, o), (v, v, v, v), (slice(2, 9), None, slice(0, 1), v),
               (slice(5, 5), o, o, o)]:
        got = basis.u.extract(*sl).cpu().numpy()
        np.testing.assert_array_equal(got, u[tuple(slice(None) if s is None else s for s in sl)])
    np.testing.assert_allclose(
        basis.compute_reference_energy(n_occ, nuclear_repulsion_energy=0.25),
        oracle.reference_energy_general(h, u, n_occ, 0.25), rtol=1e-13,
    )
    spatial = sharded.ShardedBasisSet.from_global(ctx, h, np.eye(n), u)
    np.testing.assert_allclose(
        spatial.compute_reference_energy(n_occ), oracle.reference_energy_spatial(h, u, n_occ), rtol=1e-13
    )
    half = basis.u.copy().axpby_(0.5)
    np.testing.assert_array_equal(half.gather().cpu().numpy(), 0.5 * u)
    np.testing.assert_array_equal(basis.u.gather().cpu().numpy(), u)
    combo = basis.u.copy().axpby_(2.0, half, -3.0)
    np.testing.assert_allclose(combo.gather().cpu().numpy(), 0.5 * u, rtol=1e-15, atol=1e-15)
