"""GPU parity of the reference-facing API (BasisSet / SpatialOrbitalSystem / GeneralOrbitalSystem /
ODQD) against vectors produced by the unmodified reference (tests/golden/make_golden.py) and against
the numpy oracle.  The tests read like the reference's own (tests/test_custom_system.py:26-68,
tests/test_helper.py:14-147, tests/test_one_dim_qd.py:127-186, tests/test_copy.py:6-18) and run in
both storage modes of the ``np`` hook: ``numpy`` (host arrays staged through the GPU per call) and
``quantum_systems_b200.xp`` (arrays resident in HBM)."""

import warnings

import numpy as np
import pytest

from conftest import assert_close_scaled, load_golden
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def host(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def modules():
    from quantum_systems_b200 import xp

    return {"numpy": np, "xp": xp}


@pytest.fixture(params=["numpy", "xp"])
def module(request):
    return modules()[request.param]


def spatial_basis(g, module, prefix="spas_"):
    from quantum_systems_b200 import BasisSet

    l = g[prefix + "h"].shape[0]
    dim = g[prefix + "position"].shape[0]
    bs = BasisSet(l, dim, np=module)
    bs.h = g[prefix + "h"].copy()
    bs.s = g[prefix + "s"].copy()
    bs.u = g[prefix + "u"].copy()
    bs.position = g[prefix + "position"].copy()
    bs.nuclear_repulsion_energy = float(g["spas_nuclear_repulsion_energy"])
    return bs


def check_storage(arr, module):
    if module is np:
        assert isinstance(arr, np.ndarray)
    else:
        assert isinstance(arr, torch.Tensor) and arr.is_cuda


def test_spatial_to_general_system(module):
    from quantum_systems_b200 import SpatialOrbitalSystem

    g = load_golden("systems_random")
    n = int(g["n"])
    spas = SpatialOrbitalSystem(n, spatial_basis(g, module))
    assert spas.n == n // 2 and spas.l == g["spas_h"].shape[0]
    gos = spas.construct_general_orbital_system()
    # the spatial system is untouched (deep copy, spatial_orbital_system.py:89-95)
    np.testing.assert_array_equal(host(spas.u), g["spas_u"])
    assert gos.l == 2 * spas.l and gos.n == n
    assert gos.o == slice(0, n) and gos.v == slice(n, gos.l)
    for key in ("h", "s", "u", "position", "spin_x", "spin_y", "spin_z", "spin_2", "spin_2_tb"):
        got = getattr(gos, key)
        check_storage(got, module)
        got = host(got)
        assert got.dtype == np.complex128, key
        assert_close_scaled(got, g["gos_" + key], rel=1e-13)
    # u is pure data movement: bit-exact
    np.testing.assert_array_equal(host(gos.u), g["gos_u"])
    assert gos._basis_set.anti_symmetrized_u and gos._basis_set.includes_spin


def test_fock_and_reference_energy(module):
    from quantum_systems_b200 import SpatialOrbitalSystem

    g = load_golden("systems_random")
    spas = SpatialOrbitalSystem(int(g["n"]), spatial_basis(g, module))
    gos = spas.construct_general_orbital_system()
    f = spas.construct_fock_matrix(spas.h, spas.u)
    check_storage(f, module)
    assert_close_scaled(host(f), g["spas_fock"], rel=1e-13)
    f = gos.construct_fock_matrix(gos.h, gos.u)
    assert_close_scaled(host(f), g["gos_fock"], rel=1e-13)
    np.testing.assert_allclose(spas.compute_reference_energy(), g["spas_e_ref"], rtol=1e-12)
    np.testing.assert_allclose(gos.compute_reference_energy(), g["gos_e_ref"], rtol=1e-12)
    # in-place semantics: a supplied f is zeroed, filled and returned (general_orbital_system.py:151-159)
    f_in = gos.np.zeros_like(gos.h)
    f_in += 3.0
    ret = gos.construct_fock_matrix(gos.h, gos.u, f=f_in)
    assert ret is f_in
    assert_close_scaled(host(f_in), g["gos_fock"], rel=1e-13)


@pytest.mark.parametrize("complex_u", [False, True])
@pytest.mark.parametrize("particles", [4, 6])
def test_spatial_fock_several_occupied_orbitals(module, particles, complex_u):
    """Restricted Fock matrix with n_occ >= 2 in both storage modes (spatial_orbital_system.py:150-190).  With one
    occupied orbital the [i,p,q] and [p,i,q] orders of the gathered exchange block coincide; with several they do
    not, and the host path must hand the kernel u[p,i,i,q] ordered [i,p,q]."""
    from quantum_systems_b200 import BasisSet, SpatialOrbitalSystem

    rng = np.random.default_rng(particles + 10 * complex_u)
    l = 7
    u = rng.standard_normal((l,) * 4)
    h = rng.standard_normal((l, l))
    if complex_u:
        u = u + 1j * rng.standard_normal((l,) * 4)
        h = h + 1j * rng.standard_normal((l, l))
    bs = BasisSet(l, 1, np=module)
    bs.h, bs.s, bs.u = h.copy(), np.eye(l), u.copy()
    spas = SpatialOrbitalSystem(particles, bs)
    assert spas.n == particles // 2 >= 2
    ref = oracle.construct_fock_matrix_spatial(h, u, particles // 2)
    f = spas.construct_fock_matrix(spas.h, spas.u)
    check_storage(f, module)
    assert_close_scaled(host(f), ref, rel=1e-13)
    energy = spas.compute_reference_energy()
    np.testing.assert_allclose(energy, oracle.reference_energy_spatial(h, u, particles // 2), rtol=1e-12)


def test_host_fock_real_h_complex_u_into_complex_f():
    """`f.fill(0); f += h; f += ...` of the reference accepts a real h with a complex u when the caller supplies a
    complex f (general_orbital_system.py:151-159); without f the real zeros_like(h) cannot take the sum."""
    from quantum_systems_b200 import BasisSet, SpatialOrbitalSystem

    rng = np.random.default_rng(5)
    l = 6
    u = rng.standard_normal((l,) * 4) + 1j * rng.standard_normal((l,) * 4)
    h = rng.standard_normal((l, l))
    bs = BasisSet(l, 1, np=np)
    bs.h, bs.s, bs.u = h.copy(), np.eye(l), u.copy()
    spas = SpatialOrbitalSystem(4, bs)
    f = np.full((l, l), 7.0 + 0j)
    ret = spas.construct_fock_matrix(spas.h, spas.u, f=f)
    assert ret is f
    assert_close_scaled(f, oracle.construct_fock_matrix_spatial(h.astype(complex), u, 2), rel=1e-13)
    with pytest.raises(TypeError):
        spas.construct_fock_matrix(spas.h, spas.u)


def test_change_basis_rectangular(module):
    """tests/test_custom_system.py:38-68: 5 -> 8 spatial and 10 -> 8 spin-orbitals, atol = rtol = 1e-12."""
    from quantum_systems_b200 import SpatialOrbitalSystem

    g = load_golden("systems_random")
    spas = SpatialOrbitalSystem(int(g["n"]), spatial_basis(g, module))
    gos = spas.construct_general_orbital_system()
    spin_x_before = host(gos.spin_x).copy()
    new_l = g["C_spas"].shape[1]
    spas.change_basis(g["C_spas"] if module is np else module.asarray(g["C_spas"]))
    gos.change_basis(g["C_gos"] if module is np else module.asarray(g["C_gos"]))
    assert spas.l == new_l and gos.l == new_l
    assert spas.h.shape == (new_l, new_l) and spas.u.shape == (new_l,) * 4
    assert gos.v == slice(gos.n, new_l)
    for key in ("h", "s", "u", "position"):
        check_storage(getattr(spas, key), module)
        np.testing.assert_allclose(host(getattr(spas, key)), g["spas_cb_" + key], rtol=1e-12, atol=1e-12, err_msg=key)
    for key in ("h", "s", "u", "position", "spin_2_tb"):
        np.testing.assert_allclose(host(getattr(gos, key)), g["gos_cb_" + key], rtol=1e-12, atol=1e-12, err_msg=key)
    # reference quirk kept: spin_x/y/z/spin_2 are not updated by change_basis (basis_set.py:368-372)
    np.testing.assert_array_equal(host(gos.spin_x), spin_x_before)
    assert_close_scaled(host(gos.construct_fock_matrix(gos.h, gos.u)), g["gos_cb_fock"], rel=1e-12)


def test_change_basis_explicit_c_tilde_equals_default(module):
    from quantum_systems_b200 import GeneralOrbitalSystem, RandomBasisSet

    np.random.seed(11)
    a = GeneralOrbitalSystem(2, RandomBasisSet(4, 1, np=module))
    b = a.copy_system()
    C = RandomBasisSet.get_random_elements((8, 8), np)
    a.change_basis(C)
    b.change_basis(C, C_tilde=C.conj().T)
    np.testing.assert_allclose(host(a.u), host(b.u), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(host(a.h), host(b.h), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("complex_", [False, True])
def test_static_helpers_match_oracle(module, complex_):
    """tests/test_helper.py:14-147 against the einsum / loop oracles."""
    from quantum_systems_b200 import BasisSet

    rng = np.random.default_rng(5)
    l = 6
    u = rng.standard_normal((l,) * 4) + (1j * rng.standard_normal((l,) * 4) if complex_ else 0)
    h = rng.standard_normal((l, l)) + (1j * rng.standard_normal((l, l)) if complex_ else 0)
    C = rng.standard_normal((l, l)) + (1j * rng.standard_normal((l, l)) if complex_ else 0)
    wrap = (lambda x: x) if module is np else module.asarray
    got = BasisSet.transform_two_body_elements(wrap(u), wrap(C), np=module)
    check_storage(got, module)
    ein = np.einsum("ls,kr,jq,ip,ijkl->pqrs", C, C, C.conj(), C.conj(), u, optimize=True)
    np.testing.assert_allclose(host(got), ein, atol=1e-10)
    got_h = BasisSet.transform_one_body_elements(wrap(h), wrap(C), np=module)
    np.testing.assert_allclose(host(got_h), C.conj().T @ h @ C, atol=1e-10)
    np.testing.assert_allclose(
        host(BasisSet.transform_two_body_elements(wrap(u), wrap(C), np=module, C_tilde=wrap(C.conj().T))), host(got)
    )
    spin = BasisSet.add_spin_two_body(wrap(u), np=module)
    np.testing.assert_array_equal(host(spin), oracle.add_spin_two_body(u))
    asym = BasisSet.anti_symmetrize_u(spin)
    check_storage(asym, module)
    np.testing.assert_array_equal(host(asym), oracle.add_spin_anti_symmetrize_loop(u))
    np.testing.assert_array_equal(host(BasisSet.add_spin_one_body(wrap(h), np=module)), oracle.add_spin_one_body(h))


def test_second_spin_doubling_warns(module):
    from quantum_systems_b200 import RandomBasisSet

    np.random.seed(3)
    bs = RandomBasisSet(3, 1, np=module)
    assert bs.change_to_general_orbital_basis() is bs
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        assert bs.change_to_general_orbital_basis() is None
    assert any("already been spin-doubled" in str(w.message) for w in caught)
    assert bs.l == 6


def test_assertions(module):
    from quantum_systems_b200 import BasisSet, GeneralOrbitalSystem, RandomBasisSet, SpatialOrbitalSystem

    bs = BasisSet(4, 1, np=module)
    with pytest.raises(AssertionError):
        bs.h = np.zeros((3, 3))
    with pytest.raises(AssertionError):
        bs.u = np.zeros((4, 4, 4, 3))
    with pytest.raises(AssertionError):
        bs.spin_x = np.zeros((4, 4))  # spin operators need a spin-doubled basis
    np.random.seed(4)
    with pytest.raises(AssertionError):
        SpatialOrbitalSystem(3, RandomBasisSet(4, 1, np=module))  # closed shell needs even n
    with pytest.raises(AssertionError):
        GeneralOrbitalSystem(9, RandomBasisSet(4, 1, np=module))  # n <= l
    gos = GeneralOrbitalSystem(2, RandomBasisSet(4, 1, np=module))
    with pytest.raises(AssertionError):
        SpatialOrbitalSystem(2, gos._basis_set)
    with pytest.raises(NotImplementedError):
        gos.change_to_hf_basis()


def test_copy_system_is_independent(module):
    """tests/test_copy.py:6-18."""
    from quantum_systems_b200 import GeneralOrbitalSystem, RandomBasisSet

    np.random.seed(8)
    gos = GeneralOrbitalSystem(2, RandomBasisSet(4, 2, np=module))
    clone = gos.copy_system()
    assert clone.np is gos.np
    before = host(gos.h).copy()
    clone._basis_set.h = clone._basis_set.h + 1
    np.testing.assert_array_equal(host(gos.h), before)
    np.testing.assert_allclose(host(clone.h), before + 1)


def test_change_module_round_trip():
    from quantum_systems_b200 import GeneralOrbitalSystem, RandomBasisSet, xp

    np.random.seed(9)
    gos = GeneralOrbitalSystem(2, RandomBasisSet(4, 1, np=np))
    u_host = gos.u.copy()
    gos.change_module(xp)
    assert isinstance(gos.u, torch.Tensor) and gos.u.is_cuda and gos.np is xp
    gos.change_module(np)
    assert isinstance(gos.u, np.ndarray)
    np.testing.assert_array_equal(gos.u, u_host)


@pytest.mark.parametrize("tag,kw", [
    ("ho", dict(l=6, grid_length=5, num_grid_points=101)),
    ("dw", dict(l=7, grid_length=6, num_grid_points=128, a=0.3, alpha=0.9, beta=0.1)),
])
def test_odqd_small(module, tag, kw):
    from quantum_systems_b200 import ODQD, GeneralOrbitalSystem, SpatialOrbitalSystem

    g = load_golden("odqd_small_" + tag)
    pot = ODQD.HOPotential(1.0) if tag == "ho" else ODQD.DWPotential(1.0, 5.0)
    od = ODQD(potential=pot, np=module, **kw)
    for key in ("h", "s", "spf", "position"):
        np.testing.assert_allclose(np.abs(host(getattr(od, key))), np.abs(g[key]), rtol=1e-12, atol=1e-12, err_msg=key)
    check_storage(od.u, module)
    assert host(od.u).dtype == np.float64
    np.testing.assert_allclose(np.abs(host(od.u)), np.abs(g["u"]), rtol=1e-11, atol=1e-12)
    spas = SpatialOrbitalSystem(2, od.copy_basis())
    np.testing.assert_allclose(np.abs(host(spas.construct_fock_matrix(spas.h, spas.u))), np.abs(g["spas_fock"]), atol=1e-11)
    np.testing.assert_allclose(spas.compute_reference_energy(), g["spas_e_ref"], atol=1e-11)
    gos = GeneralOrbitalSystem(2, od)
    np.testing.assert_allclose(np.abs(host(gos.u)), np.abs(g["gos_u"]), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(np.abs(host(gos.construct_fock_matrix(gos.h, gos.u))), np.abs(g["gos_fock"]), atol=1e-11)
    np.testing.assert_allclose(gos.compute_reference_energy(), g["gos_e_ref"], atol=1e-11)


def _ref_dat_system(name):
    from quantum_systems_b200 import ODQD

    if name == "odho":
        return dict(l=10, grid_length=5, num_grid_points=1001, potential=ODQD.HOPotential(1))
    if name == "oddw":
        return dict(l=10, grid_length=6, num_grid_points=1001, potential=ODQD.DWPotential(1, 5))
    if name == "odgauss":
        return dict(l=10, grid_length=20, num_grid_points=1001, potential=ODQD.GaussianPotential(1, 0, 2.5, np=np))
    return dict(l=10, grid_length=5, num_grid_points=1001, potential=ODQD.DWPotentialSmooth(a=5))


@pytest.mark.parametrize("name", ["odho", "oddw", "odgauss", "oddw_smooth"])
def test_reference_golden_files(name):
    """The reference's own fixtures, as asserted by its tests/test_one_dim_qd.py:127-186."""
    from quantum_systems_b200 import ODQD, GeneralOrbitalSystem

    g = load_golden("ref_dat_" + name)
    od = ODQD(**_ref_dat_system(name))
    u_spatial = host(od.u)
    assert np.abs(u_spatial - u_spatial.transpose(1, 0, 3, 2)).max() < 1e-8
    gos = GeneralOrbitalSystem(2, od)
    np.testing.assert_allclose(np.abs(g["dipole_moment"]), np.abs(host(gos.dipole_moment)), atol=1e-9)
    np.testing.assert_allclose(g["h"], host(gos.h), atol=1e-10)
    np.testing.assert_allclose(np.abs(g["u"]), np.abs(host(gos.u)), atol=1e-10)
    np.testing.assert_allclose(np.abs(g["spf"]), np.abs(host(gos.spf)), atol=1e-10)
    u = host(gos.u)
    assert np.abs(u + u.transpose(0, 1, 3, 2)).max() < 1e-8
    assert np.abs(u + u.transpose(1, 0, 2, 3)).max() < 1e-8
    assert np.abs(u - u.transpose(1, 0, 3, 2)).max() < 1e-8


def test_config1_odqd40_change_basis(module):
    """BASELINE.json configs[0]: ODQD(20, 10, 201) -> 40 spin-orbitals -> change_basis(orthonormal C)."""
    from quantum_systems_b200 import ODQD, GeneralOrbitalSystem

    g = load_golden("config1_odqd40_change_basis")
    od = ODQD(20, 10, 201, potential=ODQD.HOPotential(0.25), np=module)
    gos = GeneralOrbitalSystem(2, od)
    assert gos.l == 40
    # eigenvector signs are LAPACK's on the host in both implementations; pin them through the oracle
    ref = oracle.odqd_setup_basis(20, 10, 201, ODQD.HOPotential(0.25))
    ref_gos = oracle.change_to_general_orbital_basis({k: ref[k] for k in ("h", "s", "u", "position", "spf")})
    assert_close_scaled(host(gos.u), ref_gos["u"], rel=1e-12)
    gos.change_basis(g["C"] if module is np else module.asarray(g["C"]))
    idx = g["idx"]
    u = host(gos.u)
    np.testing.assert_allclose(host(gos.h), g["h"], atol=1e-10)
    np.testing.assert_allclose(u[np.ix_(idx, idx, idx, idx)], g["u_sample"], atol=1e-10)
    np.testing.assert_allclose(np.abs(u).sum(), float(g["u_abs_sum"]), rtol=1e-9)
    assert_close_scaled(u, oracle.change_basis(ref_gos, g["C"])["u"], rel=1e-12)
    # spf / bra_spf follow the basis change (basis_set.py:408-411)
    expected = oracle.change_basis(ref_gos, g["C"])
    assert_close_scaled(host(gos.spf), expected["spf"], rel=1e-12)
    assert_close_scaled(host(gos.bra_spf), expected["bra_spf"], rel=1e-12)


def test_real_storage_option_keeps_float64():
    """Deviation switch documented in DESIGN.md: real integrals may stay real through spin doubling
    (the reference always casts to complex128, basis_set.py:632-634), halving the bytes of u."""
    from quantum_systems_b200 import BasisSet, GeneralOrbitalSystem

    rng = np.random.default_rng(12)
    l = 6
    bs = BasisSet(l, 1)
    bs.cast_to_complex_on_spin_doubling = False
    u = rng.standard_normal((l,) * 4)
    bs.h = rng.standard_normal((l, l))
    bs.s = np.eye(l)
    bs.u = u
    gos = GeneralOrbitalSystem(2, bs)
    assert gos.u.dtype == torch.float64
    np.testing.assert_array_equal(host(gos.u), oracle.anti_symmetrize_u(oracle.add_spin_two_body(u)))
    C = np.linalg.qr(rng.standard_normal((2 * l, 2 * l)))[0]
    expected = oracle.transform_two_body_elements(host(gos.u), C)
    gos.change_basis(gos.np.asarray(C))
    assert gos.u.dtype == torch.float64
    assert_close_scaled(host(gos.u), expected, rel=1e-12)
