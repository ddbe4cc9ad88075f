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


def test_library_made_symmetry_is_tagged_and_the_tag_dies_with_the_data():
    """A tensor the library itself has just made exactly (anti-)symmetric -- fused spin doubling, anti_symmetrize,
    a symmetry-aware transform -- is not tested again by the next transform (no pass over u, no host sync); any
    in-place change through torch, and any new tensor, loses the tag, and then the exact test runs again."""
    from quantum_systems_b200 import ops

    n = 48
    rng = np.random.default_rng(5)
    plain = rng.standard_normal((n,) * 4)
    u = dev(plain)
    assert ops.proven_symmetry(u) == 0
    anti = ops.anti_symmetrize(u)
    assert ops.proven_symmetry(anti) == ops.ANTISYMMETRIC_LAST_PAIR
    assert ops.two_body_symmetry(anti) & 1  # the tag tells the truth
    spin = ops.add_spin_two_body(dev(plain[:24, :24, :24, :24]), anti_symmetrize=True)
    assert ops.proven_symmetry(spin) == 1 and ops.two_body_symmetry(spin) & 1
    assert ops.proven_symmetry(ops.add_spin_two_body(dev(plain[:24, :24, :24, :24]))) == 0
    assert ops.proven_symmetry(ops.add_spin_two_body(dev(plain[:24, :24, :24, :24]), True, planes=(0, 10))) == 0

    C = dev(np.linalg.qr(rng.standard_normal((n, n)))[0])
    calls = []
    real_test = ops.two_body_symmetry
    ops.two_body_symmetry = lambda *a, **k: calls.append(1) or real_test(*a, **k)
    try:
        out = ops.transform_two_body(anti, C)          # tagged input: no detection
        assert calls == [] and ops.proven_symmetry(out) == 1
        out2 = ops.transform_two_body(out, C)          # a chain of basis changes never tests
        assert calls == [] and ops.proven_symmetry(out2) == 1
        expected = oracle.transform_two_body_elements(plain - plain.transpose(0, 1, 3, 2), C.cpu().numpy())
        assert_close_scaled(out.cpu().numpy(), expected)
        assert_close_scaled(out2.cpu().numpy(), oracle.transform_two_body_elements(expected, C.cpu().numpy()), rel=1e-11)
        np.testing.assert_array_equal(out2.cpu().numpy(), -out2.cpu().numpy().transpose(0, 1, 3, 2))

        out[3, 4, 5, 6] += 1.0                         # in-place change: the tag is void, the test runs and says no
        assert ops.proven_symmetry(out) == 0
        broken = ops.transform_two_body(out, C)
        assert calls == [1] and ops.proven_symmetry(broken) == 0
        expected_broken = oracle.transform_two_body_elements(out.cpu().numpy(), C.cpu().numpy())
        assert_close_scaled(broken.cpu().numpy(), expected_broken)

        clone = anti.clone()                           # a new tensor object is tested afresh (and passes)
        assert ops.proven_symmetry(clone) == 0
        again = ops.transform_two_body(clone, C)
        assert calls == [1, 1] and ops.proven_symmetry(again) == 1
    finally:
        ops.two_body_symmetry = real_test


def test_basis_set_chain_of_basis_changes_uses_the_tag():
    """BasisSet(np=xp): the exchange symmetry of a user-supplied u is found once by the exact test; every later
    change_basis of the chain trusts the library's own mirror fill."""
    from quantum_systems_b200 import BasisSet, ops, xp

    n = 48
    rng = np.random.default_rng(9)
    u = symmetric_input(rng, n, False, "exchange")
    C = np.linalg.qr(rng.standard_normal((n, n)))[0]
    bs = BasisSet(n, 1, np=xp)
    bs.h, bs.s, bs.u = rng.standard_normal((n, n)), np.eye(n), u
    calls = []
    real_test = ops.two_body_symmetry
    ops.two_body_symmetry = lambda *a, **k: calls.append(1) or real_test(*a, **k)
    try:
        bs.change_basis(xp.asarray(C))
        bs.change_basis(xp.asarray(C))
        bs.change_basis(xp.asarray(C))
    finally:
        ops.two_body_symmetry = real_test
    assert calls == [1]
    expected = u
    for _ in range(3):
        expected = oracle.transform_two_body_elements(expected, C)
    assert_close_scaled(bs.u.cpu().numpy(), expected, rel=1e-11)
    got = bs.u.cpu().numpy()
    np.testing.assert_array_equal(got, got.transpose(1, 0, 3, 2))
    bs.u = bs.u.cpu().numpy()  # a host round trip is a new tensor: tested again
    assert ops.proven_symmetry(bs.u) == 0
