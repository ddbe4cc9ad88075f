"""CPU checks for the two-dimensional harmonic-oscillator row.

1. The C oracle (``oracle/tdho_oracle.c``) against every golden vector the reference holds for this
   path -- ``tests/dat/two_dim_quantum_dots_coulomb_elements.dat``, ``index_map.dat``, the one-body table
   (reference tests/test_two_dim_ho.py:52-90) -- and against vectors produced by running the reference
   (``tests/golden/make_golden_tdho.py``).
2. The exact rational evaluation against the oracle (it bounds the rounding of the reference's own sums).
3. The host logic of ``quantum_systems_b200.two_dim_ho`` (shell bookkeeping, one-body / position / spf,
   double wells, magnetic-field ordering) with the device kernel replaced by a stand-in defined HERE
   that calls the oracle -- the product itself has no CPU path.
"""

import numpy as np
import pytest

from conftest import load_golden
from oracle import tdho


@pytest.fixture(scope="module")
def run():
    return load_golden("tdho_reference_run")


@pytest.fixture(scope="module")
def table():
    return load_golden("tdho_reference_table")


def test_index_maps_match_reference_table(table, run):
    for p, n, m in table["index_map"]:
        assert tdho.get_indices_nm(p) == (n, m)  # reference tests/test_two_dim_ho.py:57-59
        assert tdho.get_index_p(n, m) == p  # :52-54
    for p, (n, m) in enumerate(run["indices_nm"]):
        assert tdho.get_indices_nm(p) == (n, m)


def test_one_body_elements_match_reference_table(table):
    l = int(table["one_body_index"].max()) + 1
    h = tdho.get_one_body_elements(l)
    np.testing.assert_allclose(np.diag(h)[table["one_body_index"]], table["one_body_value"], atol=1e-6, rtol=1e-6)


def test_oracle_coulomb_matches_reference_run(run):
    np.testing.assert_allclose(tdho.get_coulomb_elements(12), run["u_l12"], atol=1e-12, rtol=0)


def test_oracle_coulomb_matches_reference_run_l30_sample(run):
    n, m = tdho.quantum_numbers(30)
    rng = np.random.default_rng(30)
    pick = rng.choice(len(run["u_l30_value"]), 4000, replace=False)
    got = tdho.coulomb_sample(n, m, run["u_l30_index"][pick].astype(np.int64))
    # both are FP64 evaluations of the same alternating sums; they differ by their rounding (fast-math or not)
    np.testing.assert_allclose(got, run["u_l30_value"][pick], atol=2e-9, rtol=0)


def test_oracle_coulomb_matches_reference_golden_table(table):
    """tests/dat/two_dim_quantum_dots_coulomb_elements.dat (8 significant digits; reference tolerance 1e-6,
    tests/test_two_dim_ho.py:70-74).  A sample of the l = 36 table plus every entry below l = 20."""
    idx = table["coulomb_index"].astype(np.int64)
    val = table["coulomb_value"]
    n, m = tdho.quantum_numbers(36)
    low = np.all(idx < 20, axis=1)
    rng = np.random.default_rng(36)
    pick = np.unique(np.concatenate([np.nonzero(low)[0], rng.choice(len(val), 3000, replace=False)]))
    got = tdho.coulomb_sample(n, m, idx[pick])
    np.testing.assert_allclose(got, val[pick], atol=1e-6, rtol=1e-6)


def test_symmetry_of_oracle_elements():
    u = tdho.get_coulomb_elements(12)
    np.testing.assert_allclose(u, u.transpose(1, 0, 3, 2), atol=1e-8)  # reference tests/test_two_dim_ho.py:20-30


def test_exact_evaluation_bounds_the_oracle_rounding():
    n, m = tdho.quantum_numbers(21)
    rng = np.random.default_rng(5)
    checked = 0
    while checked < 60:
        p, q, r = rng.integers(0, 21, 3)
        match = np.nonzero(m == m[p] + m[q] - m[r])[0]
        if len(match) == 0:
            continue
        s = rng.choice(match)
        exact = tdho.coulomb_ho_exact(n[p], m[p], n[q], m[q], n[r], m[r], n[s], m[s])
        assert abs(exact - tdho.coulomb_ho(n[p], m[p], n[q], m[q], n[r], m[r], n[s], m[s])) < 1e-11
        checked += 1
    assert tdho.coulomb_ho_exact(0, 0, 0, 1, 0, 0, 0, 0) == 0.0  # m not conserved


# ------------------------------------------------------------------------------------------------
# host logic of the product, device kernel replaced by an oracle-backed stand-in
# ------------------------------------------------------------------------------------------------
@pytest.fixture()
def host_only(monkeypatch):
    pytest.importorskip("torch")
    from quantum_systems_b200 import ops

    def stand_in(n, m, scale=1.0, planes=None, device=None):
        return scale * tdho.get_coulomb_elements(len(n), n, m)

    monkeypatch.setattr(ops, "tdho_coulomb", stand_in)


def test_shell_bookkeeping_of_the_product(run, table):
    pytest.importorskip("torch")
    from quantum_systems_b200 import two_dim_ho

    for p, (n, m) in enumerate(run["indices_nm"]):
        assert two_dim_ho.get_indices_nm(p) == (n, m)
        assert two_dim_ho.get_index_p(n, m) == p
    np.testing.assert_array_equal(two_dim_ho.get_one_body_elements(36), tdho.get_one_body_elements(36))


def test_oscillator_system_host_arrays(host_only, run):
    from quantum_systems_b200 import TwoDimensionalHarmonicOscillator

    ho = TwoDimensionalHarmonicOscillator(10, 4.0, 21, omega=0.7, mass=1.3, np=np)
    for key in ("h", "u", "s", "position", "spf"):
        got = getattr(ho, key)
        assert got.dtype == run["ho_" + key].dtype
        np.testing.assert_allclose(got, run["ho_" + key], atol=1e-13, rtol=0)


def test_double_well_one_body_elements(run):
    pytest.importorskip("torch")
    from quantum_systems_b200 import two_dim_ho

    got = two_dim_ho.get_double_well_one_body_elements(12, 0.8, 1, 3, dtype=np.complex128, axis=0)
    np.testing.assert_allclose(got, run["dw_h_axis0"], atol=1e-13, rtol=0)
    got = two_dim_ho.get_double_well_one_body_elements(12, 1.0, 1, 2, dtype=np.complex128, axis=1)
    np.testing.assert_allclose(got, run["dw_h_axis1"], atol=1e-13, rtol=0)
    got = two_dim_ho.get_smooth_double_well_one_body_elements(8, 0.9, 1, a=2, b=2, dtype=np.complex128)
    np.testing.assert_allclose(got, run["smooth_dw_h"], atol=1e-13, rtol=0)
    # spectrum pinned by the reference (tests/test_two_dim_dw.py:87-111)
    h_dw = two_dim_ho.get_double_well_one_body_elements(6, 1, 1, 2, dtype=np.complex128, axis=1)
    expected = np.array([0.81129823, 1.37162083, 1.93581042, 2.21403823, 2.37162083, 2.93581042])
    np.testing.assert_allclose(np.linalg.eigvalsh(h_dw), expected, rtol=1e-7)


def test_theta_tilde_integrals_closed_forms():
    """Reference tests/test_two_dim_dw.py:18-72 (closed forms from a computer-algebra system)."""
    pytest.importorskip("torch")
    from quantum_systems_b200 import two_dim_ho

    for m_p in range(-30, 31):
        for m_q in range(-30, 31):
            d = m_q - m_p
            if abs(d) == 1:
                assert two_dim_ho.theta_1_tilde_integral(m_p, m_q) == 0
                assert two_dim_ho.theta_2_tilde_integral(m_p, m_q) == 0
                continue
            e = np.exp(1j * np.pi * d)
            one = -1j * (-d + d * e - 2j * np.exp(1j * np.pi * d / 2)) * (1 + e) / (d**2 - 1)
            two = -((1 + e) ** 2) / (d**2 - 1)
            assert abs(one - two_dim_ho.theta_1_tilde_integral(m_p, m_q)) < 1e-10
            assert abs(two - two_dim_ho.theta_2_tilde_integral(m_p, m_q)) < 1e-10


def test_magnetic_field_levels_and_system(host_only, run):
    from quantum_systems_b200 import TwoDimHarmonicOscB

    hob = TwoDimHarmonicOscB(10, 5, 21, omega_c=0.5, np=np)
    np.testing.assert_array_equal(np.stack([hob.level_n, hob.level_m], axis=1), run["hob_levels"])
    np.testing.assert_allclose(hob.level_energy, run["hob_energy"], atol=0, rtol=0)
    for key in ("h", "u", "position", "spf"):
        got = getattr(hob, key)
        assert got.dtype == run["hob_" + key].dtype
        np.testing.assert_allclose(got, run["hob_" + key], atol=1e-12, rtol=0)
    hob2 = TwoDimHarmonicOscB(7, 5, 11, omega=0.6, omega_c=1.3, np=np)
    np.testing.assert_array_equal(np.stack([hob2.level_n, hob2.level_m], axis=1), run["hob2_levels"])
    np.testing.assert_allclose(hob2.h, run["hob2_h"], atol=1e-13)
    np.testing.assert_allclose(hob2.u, run["hob2_u"], atol=1e-11)


def test_native_planner_rejects_out_of_range_shells():
    """Host-only planning entry of the C ABI: shells beyond the 14th are refused loudly."""
    import ctypes

    pytest.importorskip("torch")
    from quantum_systems_b200 import _native
    from quantum_systems_b200.build import build

    build()
    lib = _native.load()
    i64p = ctypes.POINTER(ctypes.c_int64)
    nbytes = ctypes.c_int64(0)
    n, m = tdho.quantum_numbers(105)
    assert lib.qs_tdho_coulomb_workspace_bytes(n.ctypes.data_as(i64p), m.ctypes.data_as(i64p), 105, ctypes.byref(nbytes)) == 0
    assert nbytes.value > 0
    n, m = tdho.quantum_numbers(106)
    assert lib.qs_tdho_coulomb_workspace_bytes(n.ctypes.data_as(i64p), m.ctypes.data_as(i64p), 106, ctypes.byref(nbytes)) != 0
    assert b"exceeds 13" in lib.qs_last_error()
