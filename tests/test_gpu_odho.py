"""``ODHO`` -- harmonic-oscillator grid basis with trapezoid-rule Coulomb integrals (reference
quantum_dots/one_dim/one_dim_qd.py:35-166) -- as the two-GEMM grid build with the trapezoid weights folded into
the orbital rows, against vectors from a run of the unmodified reference (tests/golden/make_golden_odho.py) and
against the oracle's restatement of the two loop nests."""

import numpy as np
import pytest

from conftest import assert_close_scaled, load_golden
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("module", ["numpy", "xp"])
def test_odho_matches_the_reference_run(tag, module):
    from quantum_systems_b200 import ODHO, xp

    g = load_golden("odho_reference_run")
    kw = {k[len(tag) + 5:]: g[k].item() for k in g if k.startswith(tag + "_arg_")}
    od = ODHO(np=np if module == "numpy" else xp, **kw)
    assert od.l == kw["l"] and od.u.shape == (kw["l"],) * 4
    for key in ("h", "s", "spf", "position", "u"):
        got = getattr(od, key)
        assert isinstance(got, np.ndarray) == (module == "numpy"), key
        got = host(got)
        assert got.dtype == g[f"{tag}_{key}"].dtype, key
        # the golden u went through numba's sequential dot products, ours through tensor-core GEMMs: 1e-12
        assert_close_scaled(got, g[f"{tag}_{key}"], rel=1e-12)
    np.testing.assert_allclose(od.eigen_energies, g[f"{tag}_eigen_energies"], rtol=1e-15)
    assert abs(kw["omega"] * 0.5 - host(od.h)[0, 0]) == 0


def test_odho_against_the_oracle_and_through_a_general_orbital_system():
    from quantum_systems_b200 import ODHO, GeneralOrbitalSystem

    kw = dict(l=16, grid_length=9.0, num_grid_points=257, omega=0.7, a=0.2, alpha=1.3)
    od = ODHO(**kw)
    ref = oracle.odho_setup_basis(**kw)
    assert_close_scaled(host(od.u), ref["u"], rel=1e-12)
    assert_close_scaled(host(od.spf), ref["spf"], rel=1e-13)
    assert_close_scaled(host(od.position), ref["position"], rel=1e-13)
    # symmetries of a real interaction on real orbitals (reference tests/test_one_dim_qd.py:146-186 for ODQD)
    u = host(od.u)
    np.testing.assert_allclose(u, u.transpose(1, 0, 3, 2), atol=1e-13)
    np.testing.assert_allclose(u, u.transpose(2, 3, 0, 1), atol=1e-13)
    gos = GeneralOrbitalSystem(2, od)
    ref_gos = oracle.change_to_general_orbital_basis({k: ref[k] for k in ("h", "s", "u", "position")})
    assert_close_scaled(host(gos.u), ref_gos["u"], rel=1e-12)
    f = gos.construct_fock_matrix(gos.h, gos.u)
    assert_close_scaled(host(f), oracle.construct_fock_matrix_general(ref_gos["h"], ref_gos["u"], 2), rel=1e-12)
