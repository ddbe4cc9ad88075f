"""Pin the numpy oracle to the reference: every oracle function against vectors produced by the
unmodified reference (tests/golden/make_golden.py) and against the reference's own golden files
(tests/dat/od*_*.npy, re-packed as ref_dat_*.npz).  CPU only."""

import numpy as np
import pytest

from conftest import assert_close_scaled, load_golden
from oracle import qs_oracle as oracle


def test_transform_square_complex():
    g = load_golden("transform_square_complex")
    np.testing.assert_allclose(oracle.transform_two_body_elements(g["u"], g["C"]), g["u_out"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(oracle.transform_one_body_elements(g["h"], g["C"]), g["h_out"], rtol=1e-13, atol=1e-13)
    # the reference's own independent check (tests/test_helper.py:43-57): einsum, atol 1e-10
    C = g["C"]
    ein = np.einsum("ls,kr,jq,ip,ijkl->pqrs", C, C, C.conj(), C.conj(), g["u"], optimize=True)
    np.testing.assert_allclose(ein, oracle.transform_two_body_elements(g["u"], C), atol=1e-10)
    np.testing.assert_allclose(
        oracle.transform_two_body_elements(g["u"], C, C.conj().T), oracle.transform_two_body_elements(g["u"], C)
    )


@pytest.mark.parametrize("name", ["transform_rect_real", "transform_biorthogonal", "transform_real_u_complex_C"])
def test_transform_variants(name):
    g = load_golden(name)
    out = oracle.transform_two_body_elements(g["u"], g["C"], g.get("C_tilde"))
    assert out.dtype == g["u_out"].dtype and out.shape == g["u_out"].shape
    np.testing.assert_allclose(out, g["u_out"], rtol=1e-12, atol=1e-12)
    if "h" in g:
        np.testing.assert_allclose(
            oracle.transform_one_body_elements(g["h"], g["C"], g.get("C_tilde")), g["h_out"], rtol=1e-12, atol=1e-12
        )


@pytest.mark.parametrize("name", ["add_spin_antisym_real", "add_spin_antisym_complex"])
def test_add_spin_antisym(name):
    g = load_golden(name)
    spin = oracle.add_spin_two_body(g["u"])
    np.testing.assert_array_equal(spin, g["u_spin"])
    np.testing.assert_array_equal(oracle.anti_symmetrize_u(spin), g["u_as"])
    np.testing.assert_array_equal(oracle.add_spin_anti_symmetrize_loop(g["u"]), g["u_as"])
    if "h" in g:
        np.testing.assert_array_equal(oracle.add_spin_one_body(g["h"]), g["h_spin"])
    # symmetry properties of tests/test_helper.py:136-147 (input has u_pqrs = u_qpsr)
    if name.endswith("real"):
        u = g["u_as"]
        np.testing.assert_allclose(u, -u.transpose(0, 1, 3, 2), atol=1e-10)
        np.testing.assert_allclose(u, -u.transpose(1, 0, 2, 3), atol=1e-10)
        np.testing.assert_allclose(u, u.transpose(1, 0, 3, 2), atol=1e-10)


def test_spin_delta():
    for p in range(20):
        for q in range(20):
            assert oracle.spin_delta(p, q) == ((p % 2) == (q % 2))


def test_systems_random():
    g = load_golden("systems_random")
    n = int(g["n"])
    spas = {k: g["spas_" + k] for k in ("h", "s", "u", "position")}
    gos = oracle.change_to_general_orbital_basis(spas)
    for key in ("h", "s", "u", "position", "spin_x", "spin_y", "spin_z", "spin_2", "spin_2_tb"):
        np.testing.assert_allclose(gos[key], g["gos_" + key], rtol=1e-13, atol=1e-13, err_msg=key)
        assert gos[key].dtype == np.complex128
    # Fock matrices and reference energies (the reference's tests never call these: golden only)
    np.testing.assert_allclose(
        oracle.construct_fock_matrix_spatial(spas["h"], spas["u"], n // 2), g["spas_fock"], rtol=1e-13, atol=1e-13
    )
    np.testing.assert_allclose(
        oracle.construct_fock_matrix_general(gos["h"], gos["u"], n), g["gos_fock"], rtol=1e-13, atol=1e-13
    )
    e_n = float(g["spas_nuclear_repulsion_energy"])
    np.testing.assert_allclose(oracle.reference_energy_spatial(spas["h"], spas["u"], n // 2, e_n), g["spas_e_ref"])
    np.testing.assert_allclose(oracle.reference_energy_general(gos["h"], gos["u"], n, e_n), g["gos_e_ref"])
    # rectangular change of basis (tests/test_custom_system.py:38-68, atol = rtol = 1e-12)
    spas_cb = oracle.change_basis(spas, g["C_spas"])
    gos_cb = oracle.change_basis(gos, g["C_gos"])
    for key in ("h", "s", "u", "position"):
        np.testing.assert_allclose(spas_cb[key], g["spas_cb_" + key], rtol=1e-12, atol=1e-12, err_msg=key)
    for key in ("h", "s", "u", "position", "spin_2_tb", "spin_x", "spin_y", "spin_z", "spin_2"):
        np.testing.assert_allclose(gos_cb[key], g["gos_cb_" + key], rtol=1e-12, atol=1e-12, err_msg=key)
    # the reference leaves spin_x/y/z/spin_2 untouched by change_basis (basis_set.py:368-372)
    np.testing.assert_array_equal(g["gos_cb_spin_x"], g["gos_spin_x"])
    np.testing.assert_allclose(
        oracle.construct_fock_matrix_general(gos_cb["h"], gos_cb["u"], n), g["gos_cb_fock"], rtol=1e-12, atol=1e-12
    )


def _potential(tag):
    from quantum_systems_b200.potentials import DWPotential, HOPotential

    return HOPotential(1.0) if tag == "ho" else DWPotential(1.0, 5.0)


@pytest.mark.parametrize("tag,kw", [
    ("ho", dict(l=6, grid_length=5, num_grid_points=101)),
    ("dw", dict(l=7, grid_length=6, num_grid_points=128, a=0.3, alpha=0.9, beta=0.1)),
])
def test_odqd_small(tag, kw):
    g = load_golden("odqd_small_" + tag)
    od = oracle.odqd_setup_basis(potential=_potential(tag), **kw)
    for key in ("h", "s", "spf", "position", "eigen_energies", "grid"):
        np.testing.assert_allclose(np.abs(od[key]), np.abs(g[key]), rtol=1e-12, atol=1e-12, err_msg=key)
    # eigenvector signs are LAPACK's; u is even in every orbital pair product only up to that sign
    np.testing.assert_allclose(np.abs(od["u"]), np.abs(g["u"]), rtol=1e-11, atol=1e-12)
    gos = oracle.change_to_general_orbital_basis({k: od[k] for k in ("h", "s", "u", "position", "spf")})
    np.testing.assert_allclose(np.abs(gos["u"]), np.abs(g["gos_u"]), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(
        np.abs(oracle.construct_fock_matrix_general(gos["h"], gos["u"], 2)), np.abs(g["gos_fock"]), atol=1e-11
    )
    np.testing.assert_allclose(
        np.abs(oracle.construct_fock_matrix_spatial(od["h"], od["u"], 1)), np.abs(g["spas_fock"]), atol=1e-11
    )
    np.testing.assert_allclose(oracle.reference_energy_general(gos["h"], gos["u"], 2), g["gos_e_ref"], atol=1e-11)
    np.testing.assert_allclose(oracle.reference_energy_spatial(od["h"], od["u"], 1), g["spas_e_ref"], atol=1e-11)


def _ref_dat_system(name):
    from quantum_systems_b200 import potentials as pot

    if name == "odho":
        return dict(l=10, grid_length=5, num_grid_points=1001, potential=pot.HOPotential(1))
    if name == "oddw":
        return dict(l=10, grid_length=6, num_grid_points=1001, potential=pot.DWPotential(1, 5))
    if name == "odgauss":
        return dict(l=10, grid_length=20, num_grid_points=1001, potential=pot.GaussianPotential(1, 0, 2.5, np=np))
    return dict(l=10, grid_length=5, num_grid_points=1001, potential=pot.DWPotentialSmooth(a=5))


@pytest.mark.parametrize("name", ["odho", "oddw", "odgauss", "oddw_smooth"])
def test_reference_golden_files(name):
    """The reference's own fixtures for this path: tests/test_one_dim_qd.py:127-143 (abs compare, 1e-10)."""
    g = load_golden("ref_dat_" + name)
    od = oracle.odqd_setup_basis(**_ref_dat_system(name))
    gos = oracle.change_to_general_orbital_basis({k: od[k] for k in ("h", "s", "u", "position", "spf")})
    np.testing.assert_allclose(np.abs(g["dipole_moment"]), np.abs(gos["position"]), atol=1e-9)
    np.testing.assert_allclose(g["h"], gos["h"], atol=1e-10)
    np.testing.assert_allclose(np.abs(g["u"]), np.abs(gos["u"]), atol=1e-10)
    np.testing.assert_allclose(np.abs(g["spf"]), np.abs(gos["spf"]), atol=1e-10)
    # symmetry tests of tests/test_one_dim_qd.py:146-186
    u = gos["u"]
    assert np.abs(u + u.transpose(0, 1, 3, 2)).max() < 1e-8
    assert np.abs(u + u.transpose(1, 0, 2, 3)).max() < 1e-8
    assert np.abs(u - u.transpose(1, 0, 3, 2)).max() < 1e-8
    assert np.abs(od["u"] - od["u"].transpose(1, 0, 3, 2)).max() < 1e-8


def test_config1_sample():
    """BASELINE.json configs[0] at its real size: ODQD(20, 10, 201) -> 40 spin-orbitals -> change_basis."""
    from quantum_systems_b200.potentials import HOPotential

    g = load_golden("config1_odqd40_change_basis")
    od = oracle.odqd_setup_basis(20, 10, 201, HOPotential(0.25))
    gos = oracle.change_to_general_orbital_basis({k: od[k] for k in ("h", "s", "u", "position", "spf")})
    out = oracle.change_basis(gos, g["C"])
    idx = g["idx"]
    # sign freedom of the eigenvectors cancels only in even products: compare magnitudes of invariants
    np.testing.assert_allclose(np.abs(out["u"]).sum(), float(g["u_abs_sum"]), rtol=1e-9)
    np.testing.assert_allclose(np.abs(out["u"]).max(), float(g["u_max"]), rtol=1e-9)
    np.testing.assert_allclose(out["h"], g["h"], atol=1e-10)
    np.testing.assert_allclose(out["u"][np.ix_(idx, idx, idx, idx)], g["u_sample"], atol=1e-10)


# --------------------------------------------------------------------------------------------
# sinc-DVR (reference-run vectors; the reference has no test of this class)
# --------------------------------------------------------------------------------------------


def test_sinc_dvr_oracle_matches_reference_run():
    from quantum_systems_b200 import potentials

    g = load_golden("sinc_dvr_reference_run")
    for repr_ in ("2d", "4d"):
        basis = oracle.sinc_dvr_setup_basis(
            14, 6.0, potentials.DWPotential(1.0, 2.0), a=0.3, alpha=0.9, beta=0.1, u_repr=repr_
        )
        for key in ("h", "s", "u", "spf", "position"):
            assert basis[key].dtype == g[f"{repr_}_{key}"].dtype
            np.testing.assert_allclose(basis[key], g[f"{repr_}_{key}"], atol=1e-14, rtol=0)
    w, C, Ct = g["2d_u"], g["C"], g["Ct"]
    np.testing.assert_allclose(oracle.sinc_dvr_transform_two_body_elements(w, C), g["tb_default"], atol=1e-12)
    np.testing.assert_allclose(oracle.sinc_dvr_transform_two_body_elements(w, C, Ct), g["tb_biorth"], atol=1e-12)
    np.testing.assert_allclose(
        oracle.sinc_dvr_transform_two_body_elements(w, C, Ct, anti_symmetrize=True), g["tb_antisym"], atol=1e-12
    )
    np.testing.assert_allclose(oracle.sinc_dvr_transform_two_body_elements(w, g["Cr"]), g["tb_real_C"], atol=1e-12)
    # the structured transform equals the dense four-index transform of the 4-D representation
    np.testing.assert_allclose(oracle.transform_two_body_elements(g["4d_u"], C, Ct), g["tb_biorth"], atol=1e-11)
    np.testing.assert_allclose(g["tb_dense_4d"], g["tb_biorth"], atol=1e-11)
    np.testing.assert_array_equal(oracle.sinc_dvr_add_spin_two_body(np.arange(4.0).reshape(2, 2))[:2, :2], np.zeros((2, 2)))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_odho_reference_run(tag):
    """Oracle restatement of the ODHO trapezoid path against a run of the unmodified reference
    (tests/golden/make_golden_odho.py; one_dim_qd.py:35-166)."""
    g = load_golden("odho_reference_run")
    kw = {k[len(tag) + 5:]: g[k].item() for k in g if k.startswith(tag + "_arg_")}
    od = oracle.odho_setup_basis(**kw)
    for key in ("h", "s", "spf", "position", "grid", "eigen_energies"):
        assert od[key].dtype == g[f"{tag}_{key}"].dtype, key
        assert_close_scaled(od[key], g[f"{tag}_{key}"], rel=1e-14)
    assert od["u"].dtype == np.complex128
    assert_close_scaled(od["u"], g[f"{tag}_u"], rel=1e-13)
