"""Symmetry-aware four-index transform: exact detection of u[p,q,r,s] = -u[p,q,s,r] and u[p,q,r,s] = u[q,p,s,r],
quarter steps 2-4 restricted to the tiles that hold a pair r < s (r <= s), mirror fill -- against the numpy oracle
(which always does the four full steps) to 1e-12, and against the plain GPU path."""

import numpy as np
import pytest

from conftest import assert_close_scaled
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if complex_ else x


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def symmetric_input(rng, n, complex_, kind):
    u = rand(rng, (n,) * 4, complex_)
    if kind == "antisym":
        return u - u.transpose(0, 1, 3, 2)
    if kind == "exchange":
        return 0.5 * (u + u.transpose(1, 0, 3, 2))
    if kind == "both":  # anti-symmetrised physical interaction: both properties at once
        u = 0.5 * (u + u.transpose(1, 0, 3, 2))
        return u - u.transpose(0, 1, 3, 2)
    return u


@pytest.mark.parametrize("kind,expected", [("none", 0), ("antisym", 1), ("exchange", 2), ("both", 3)])
@pytest.mark.parametrize("complex_", [False, True])
@pytest.mark.parametrize("n", [3, 31, 32, 50])
def test_symmetry_detection_is_exact(n, complex_, kind, expected):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(n)
    u = symmetric_input(rng, n, complex_, kind)
    assert ops.two_body_symmetry(dev(u)) == expected
    if expected:
        # one element off by one ulp breaks every symmetry it takes part in
        broken = u.copy()
        idx = (n - 1, 0, n // 2, n - 1) if n > 2 else (0, 0, 0, 1)
        broken[idx] = np.nextafter(broken[idx].real, np.inf) + 1j * broken[idx].imag if complex_ else np.nextafter(broken[idx], np.inf)
        assert ops.two_body_symmetry(dev(broken)) == 0
    assert ops.two_body_symmetry(torch.zeros((n,) * 4, dtype=torch.float64, device="cuda")) == 3


@pytest.mark.parametrize("kind", ["antisym", "exchange", "both"])
@pytest.mark.parametrize("u_complex,c_complex,biorth", [(False, False, False), (True, False, False), (True, True, True),
                                                        (False, True, False)])
@pytest.mark.parametrize("n,m", [(48, 48), (50, 64), (64, 50), (96, 96), (70, 33)])
def test_symmetric_transform_matches_oracle(n, m, u_complex, c_complex, biorth, kind):
    from quantum_systems_b200 import ops

    if n == 96 and (u_complex or c_complex):
        pytest.skip("n = 96 is checked against the real oracle only (the complex one takes ~3 s per case)")
    rng = np.random.default_rng(7 * n + m)
    u = symmetric_input(rng, n, u_complex, kind)
    C = rand(rng, (n, m), c_complex)
    Ct = rand(rng, (m, n), c_complex) if biorth else None
    expected = oracle.transform_two_body_elements(u, C, Ct)
    u_dev, C_dev, Ct_dev = dev(u), dev(C), None if Ct is None else dev(Ct)
    flags = ops.two_body_symmetry(u_dev)
    assert flags == {"antisym": 1, "exchange": 2, "both": 3}[kind]
    got = ops.transform_two_body(u_dev, C_dev, Ct_dev)  # detects and exploits the symmetry (n, m >= 48) ...
    assert_close_scaled(got.cpu().numpy(), expected)
    plain = ops.transform_two_body(u_dev, C_dev, Ct_dev, symmetry=0)  # ... unless told not to
    assert_close_scaled(plain.cpu().numpy(), expected)
    for forced in (1, 2):
        if flags & forced:
            out = ops.transform_two_body(u_dev, C_dev, Ct_dev, symmetry=forced)
            assert_close_scaled(out.cpu().numpy(), expected)
            # the completed result carries the symmetry exactly (the mirror image is a copy, not a recomputation)
            assert ops.two_body_symmetry(out) & forced


def test_symmetry_is_preserved_through_a_gos_change_basis():
    """GeneralOrbitalSystem path: spin doubling + anti-symmetrisation gives an exactly antisymmetric u, the basis
    change keeps it exactly antisymmetric, and the Fock matrix built from it matches the oracle."""
    from quantum_systems_b200 import BasisSet, GeneralOrbitalSystem, ops

    l = 30
    rng = np.random.default_rng(30)
    u = rng.standard_normal((l,) * 4)
    u = 0.5 * (u + u.transpose(1, 0, 3, 2))
    h = rng.standard_normal((l, l))
    bs = BasisSet(l, 1)
    bs.h, bs.u, bs.s = h, u, np.eye(l)
    gos = GeneralOrbitalSystem(4, bs)
    assert ops.two_body_symmetry(gos.u) == 3
    C = np.linalg.qr(rng.standard_normal((2 * l, 2 * l)))[0]
    ref = oracle.change_basis(oracle.change_to_general_orbital_basis({"h": h, "s": np.eye(l), "u": u}), C)
    gos.change_basis(gos.np.asarray(C))
    assert_close_scaled(gos.u.cpu().numpy(), ref["u"])
    assert ops.two_body_symmetry(gos.u) & 1
    f = gos.construct_fock_matrix(gos.h, gos.u)
    assert_close_scaled(f.cpu().numpy(), oracle.construct_fock_matrix_general(ref["h"], ref["u"], 4))
