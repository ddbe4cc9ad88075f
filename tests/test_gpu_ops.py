"""GPU parity of the tensor-level operators (through the C ABI) against the numpy oracle."""

import numpy as np
import pytest

from conftest import assert_close_scaled, load_golden
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.cpu().numpy()


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    if complex_:
        x = x + 1j * rng.standard_normal(shape)
    return x


@pytest.mark.parametrize("u_complex", [False, True])
@pytest.mark.parametrize("c_complex", [False, True])
@pytest.mark.parametrize("n,m", [(10, 10), (9, 6), (7, 11), (20, 18), (40, 40), (33, 70)])
def test_transform_two_body_matches_oracle(n, m, u_complex, c_complex):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(1000 * n + 10 * m + 2 * u_complex + c_complex)
    u = rand(rng, (n, n, n, n), u_complex)
    C = rand(rng, (n, m), c_complex)
    expected = oracle.transform_two_body_elements(u, C)
    got = host(ops.transform_two_body(dev(u), dev(C)))
    assert got.dtype == expected.dtype
    assert_close_scaled(got, expected)


@pytest.mark.parametrize("c_complex", [False, True])
def test_transform_two_body_biorthogonal(c_complex):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(7)
    n = 24
    u = rand(rng, (n, n, n, n), True)
    C = rand(rng, (n, n), c_complex) + 3 * np.eye(n)
    Ct = np.linalg.inv(C)
    expected = oracle.transform_two_body_elements(u, C, Ct)
    got = host(ops.transform_two_body(dev(u), dev(C), dev(Ct)))
    assert_close_scaled(got, expected)


@pytest.mark.parametrize("name", ["transform_square_complex", "transform_rect_real", "transform_biorthogonal",
                                  "transform_real_u_complex_C"])
def test_transform_golden(name):
    from quantum_systems_b200 import ops

    g = load_golden(name)
    Ct = dev(g["C_tilde"]) if "C_tilde" in g else None
    got = host(ops.transform_two_body(dev(g["u"]), dev(g["C"]), Ct))
    assert_close_scaled(got, g["u_out"])
    if "h" in g:
        got_h = host(ops.transform_one_body(dev(g["h"]), dev(g["C"]), Ct))
        assert_close_scaled(got_h, g["h_out"])


@pytest.mark.parametrize("in_complex,out_complex", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("l", [1, 5, 16, 33, 40])
@pytest.mark.parametrize("antisym", [False, True])
def test_add_spin_two_body(l, in_complex, out_complex, antisym):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(l)
    u = rand(rng, (l, l, l, l), in_complex)
    expected = oracle.add_spin_two_body(u)
    if antisym:
        expected = oracle.anti_symmetrize_u(expected)
    got = host(ops.add_spin_two_body(dev(u), anti_symmetrize=antisym,
                                     out_dtype=torch.complex128 if out_complex else torch.float64))
    np.testing.assert_array_equal(got, expected.astype(got.dtype))  # pure data movement: bit-exact


@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("n", [2, 10, 37, 64])
def test_anti_symmetrize(n, complex_):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(n)
    u = rand(rng, (n, n, n, n), complex_)
    got = host(ops.anti_symmetrize(dev(u)))
    np.testing.assert_array_equal(got, oracle.anti_symmetrize_u(u))


@pytest.mark.parametrize("h_complex,u_complex", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("n,n_occ", [(8, 2), (20, 6), (40, 40), (50, 0), (70, 33)])
def test_fock(n, n_occ, h_complex, u_complex):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(n + n_occ)
    h = rand(rng, (n, n), h_complex)
    u = rand(rng, (n, n, n, n), u_complex)
    got = host(ops.fock_general(dev(h), dev(u), n_occ))
    assert_close_scaled(got, oracle.construct_fock_matrix_general(h, u, n_occ), rel=1e-14 * max(n_occ, 1))
    got = host(ops.fock_spatial(dev(h), dev(u), n_occ))
    assert_close_scaled(got, oracle.construct_fock_matrix_spatial(h, u, n_occ), rel=1e-14 * max(n_occ, 1))
    # in-place semantics: the caller's f is overwritten and returned
    f = torch.full((n, n), 7.0, dtype=dev(h).dtype, device="cuda")
    ret = ops.fock_general(dev(h), dev(u), n_occ, f=f)
    assert ret is f
    assert_close_scaled(host(f), oracle.construct_fock_matrix_general(h, u, n_occ), rel=1e-14 * max(n_occ, 1))


@pytest.mark.parametrize("l,G", [(6, 101), (7, 128), (20, 201), (10, 1001)])
def test_odqd_coulomb(l, G):
    from quantum_systems_b200 import ops

    grid, eps, C = oracle.odqd_orbitals(l, 5.0, G, lambda x: 0.5 * x**2)
    expected = oracle.odqd_coulomb_elements(C, grid, 1.0, 0.25)
    got = host(ops.odqd_coulomb(dev(C), dev(grid[1:-1]), 1.0, 0.25))
    assert_close_scaled(got, expected)


def test_cpu_tensor_is_rejected():
    from quantum_systems_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.anti_symmetrize(torch.zeros((2, 2, 2, 2), dtype=torch.float64))
