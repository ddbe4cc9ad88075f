"""Edge cases of the hot-path operators on the GPU: the smallest extents (n = 1, 2, 3; one new orbital), odd sizes
straddling every tile boundary of the quarter GEMM (rows 128, columns 8 / 64, k 8 / 16), non-contiguous and
non-float64 inputs (the reference's ODQD hands out a permuted einsum view; random test fixtures are float/int
mixes), empty occupied spaces, and loud failures for malformed arguments -- all against the numpy oracle."""

import numpy as np
import pytest

from conftest import assert_close_scaled
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.cpu().numpy()


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if complex_ else x


@pytest.mark.parametrize("u_complex,c_complex", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("n,m", [(1, 1), (2, 2), (3, 3), (1, 4), (5, 1), (2, 9), (8, 8), (15, 17), (16, 16), (17, 15),
                                 (65, 7), (7, 65)])
def test_transform_at_tile_boundaries(n, m, u_complex, c_complex):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(97 * n + m)
    u, C = rand(rng, (n,) * 4, u_complex), rand(rng, (n, m), c_complex)
    Ct = rand(rng, (m, n), c_complex)
    for bra in (None, Ct):
        expected = oracle.transform_two_body_elements(u, C, bra)
        got = host(ops.transform_two_body(dev(u), dev(C), None if bra is None else dev(bra)))
        assert got.shape == (m,) * 4 and got.dtype == expected.dtype
        assert_close_scaled(got, expected)
    h = rand(rng, (n, n), u_complex)
    assert_close_scaled(host(ops.transform_one_body(dev(h), dev(C), dev(Ct))), oracle.transform_one_body_elements(h, C, Ct))


def test_non_contiguous_and_non_float64_inputs():
    from quantum_systems_b200 import BasisSet, ops, xp

    rng = np.random.default_rng(5)
    n = 12
    base = rng.standard_normal((n,) * 4)
    view = base.transpose(0, 2, 1, 3)  # what np.einsum(..., optimize=True) returns in the reference's ODQD
    assert not view.flags.c_contiguous
    C = np.linalg.qr(rng.standard_normal((n, n)))[0]
    expected = oracle.transform_two_body_elements(view, C)
    assert_close_scaled(host(ops.transform_two_body(torch.from_numpy(base).cuda().permute(0, 2, 1, 3), dev(C))), expected)
    for module in (np, xp):
        bs = BasisSet(n, 1, np=module)
        bs.u = view                                   # non-contiguous ndarray
        bs.h = np.arange(n * n).reshape(n, n)         # integers
        bs.s = np.eye(n, dtype=np.float32)            # single precision
        bs.change_basis(C.astype(np.float64))
        got_u = bs.u if isinstance(bs.u, np.ndarray) else host(bs.u)
        got_h = bs.h if isinstance(bs.h, np.ndarray) else host(bs.h)
        assert_close_scaled(got_u, expected)
        assert_close_scaled(got_h, oracle.transform_one_body_elements(np.arange(n * n).reshape(n, n).astype(float), C))


@pytest.mark.parametrize("l", [1, 2, 3])
def test_smallest_spin_doubling_and_fock(l):
    from quantum_systems_b200 import BasisSet, GeneralOrbitalSystem, SpatialOrbitalSystem

    rng = np.random.default_rng(l)
    h, u = rand(rng, (l, l), True), rand(rng, (l,) * 4, True)
    bs = BasisSet(l, 1, np=np)
    bs.h, bs.u, bs.s = h.copy(), u.copy(), np.eye(l)
    spas = SpatialOrbitalSystem(0, bs.copy_basis())   # no particles: f = h, E = 0
    np.testing.assert_array_equal(spas.construct_fock_matrix(spas.h, spas.u), h)
    assert spas.compute_reference_energy() == 0
    gos = GeneralOrbitalSystem(2 * l, bs)             # every spin-orbital occupied
    ref = oracle.change_to_general_orbital_basis({"h": h, "s": np.eye(l), "u": u})
    np.testing.assert_array_equal(gos.u, ref["u"])
    assert_close_scaled(gos.construct_fock_matrix(gos.h, gos.u), oracle.construct_fock_matrix_general(ref["h"], ref["u"], 2 * l))
    np.testing.assert_allclose(gos.compute_reference_energy(), oracle.reference_energy_general(ref["h"], ref["u"], 2 * l),
                               rtol=1e-13)
    f = np.full((2 * l, 2 * l), 3.0 + 0j)
    assert gos.construct_fock_matrix(gos.h, gos.u, f=f) is f  # host f filled in place, like the reference


def test_empty_plane_ranges_are_no_ops():
    from quantum_systems_b200 import ops

    u = dev(np.random.default_rng(0).standard_normal((4,) * 4))
    assert ops.add_spin_two_body(u, planes=(3, 3)).shape == (0, 8, 8, 8)
    assert ops.anti_symmetrize(u[:0].contiguous()).shape == (0, 4, 4, 4)
    assert ops.extract_block(u, slice(2, 2)).shape == (0, 4, 4, 4)
    assert ops.scale_add(u[:0].contiguous(), 2.0).numel() == 0


def test_malformed_arguments_fail_loudly():
    from quantum_systems_b200 import BasisSet, ops

    u = dev(np.zeros((4,) * 4))
    with pytest.raises(ValueError):
        ops.transform_two_body(u, dev(np.eye(5)))                    # C does not match u
    with pytest.raises(ValueError):
        ops.transform_two_body(u, dev(np.eye(4)), dev(np.zeros((3, 4))))  # C_tilde of the wrong shape
    with pytest.raises(TypeError):
        ops.transform_two_body(u.to(torch.float32), dev(np.eye(4)).to(torch.float32))
    with pytest.raises(ValueError):
        ops.add_spin_two_body(dev(np.zeros((2, 2, 2, 3))))
    with pytest.raises(RuntimeError, match="n_occ"):
        ops.fock_general(dev(np.eye(4)), u, 5)
    bs = BasisSet(4, 1)
    with pytest.raises(AssertionError):
        bs.u = np.zeros((4, 4, 4, 5))                                # reference basis_set.py:101-105
    with pytest.raises(AssertionError):
        bs.h = np.zeros((3, 3))
