"""The closed form used for full-size parity (tests/closed_form.py) against the oracle's explicit pipeline
ODQD build -> add_spin -> anti-symmetrise -> change_basis at a size the oracle handles in milliseconds."""

import numpy as np
import pytest

from oracle import qs_oracle as oracle

import closed_form


@pytest.mark.parametrize("biorthogonal", [False, True])
@pytest.mark.parametrize("anti", [True, False])
def test_closed_form_equals_explicit_pipeline(biorthogonal, anti):
    l, G = 5, 41
    grid, eps, L = oracle.odqd_orbitals(l, 4.0, G, lambda x: 0.5 * x**2)
    u = oracle.odqd_coulomb_elements(L, grid, 1.0, 0.25)
    spin = oracle.add_spin_two_body(np.ascontiguousarray(u))
    if anti:
        spin = oracle.anti_symmetrize_u(spin)
    rng = np.random.default_rng(7)
    n = 2 * l
    if biorthogonal:
        C = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + 3 * np.eye(n)
        Ct = np.linalg.inv(C)
    else:
        C, Ct = np.linalg.qr(rng.standard_normal((n, n)))[0], None
    expected = oracle.transform_two_body_elements(spin, C, Ct)
    W = closed_form.shielded_coulomb(grid[1:-1], 1.0, 0.25)
    form = closed_form.SpinDoubledClosedForm(L, W, C, Ct, anti_symmetrize=anti)
    scale = np.abs(expected).max()
    assert np.abs(form.dense() - expected).max() <= 1e-12 * scale
    assert np.abs(form.plane(3, 7) - expected[3, 7]).max() <= 1e-12 * scale
    idx = rng.integers(0, n, size=(4, 500))
    assert np.abs(form.elements(*idx) - expected[idx[0], idx[1], idx[2], idx[3]]).max() <= 1e-12 * scale
    ref_f = oracle.construct_fock_matrix_general(np.zeros((n, n), dtype=expected.dtype), expected, 3)
    assert np.abs(form.fock_two_body(3) - ref_f).max() <= 1e-12 * np.abs(ref_f).max()
    # two successive basis changes compose into the net one: C1 C2 on the right, C~2 C~1 on the left
    second = oracle.transform_two_body_elements(expected, C, Ct)
    Ct1 = C.conj().T if Ct is None else Ct
    net = closed_form.SpinDoubledClosedForm(L, W, C @ C, Ct1 @ Ct1, anti_symmetrize=anti)
    assert np.abs(net.dense() - second).max() <= 1e-11 * np.abs(second).max()


def test_check_shard_flags_a_wrong_element():
    import torch

    l, G = 4, 31
    grid, eps, L = oracle.odqd_orbitals(l, 4.0, G, lambda x: 0.5 * x**2)
    W = closed_form.shielded_coulomb(grid[1:-1], 1.0, 0.25)
    C = np.linalg.qr(np.random.default_rng(1).standard_normal((2 * l, 2 * l)))[0]
    form = closed_form.SpinDoubledClosedForm(L, W, C)
    full = form.dense()
    slab = torch.from_numpy(np.ascontiguousarray(full[2:5]))
    errs = closed_form.check_shard(closed_form.TorchSlab(slab), 2, form, np.random.default_rng(0), samples=300)
    assert max(errs[:3]) <= 1e-13 * errs[3]
    slab[0, 0, 1, 2] += 1e-6
    errs = closed_form.check_shard(closed_form.TorchSlab(slab), 2, form, np.random.default_rng(0), samples=300)
    assert errs[0] > 1e-9 * errs[3]
