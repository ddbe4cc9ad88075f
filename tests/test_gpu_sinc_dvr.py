"""GPU parity of ``ODSincDVR``: the structured transform of a diagonal two-body
operator (two chained DMMA GEMMs, ``qs_transform_two_body_diagonal``) and the overridden ``BasisSet`` hooks,
against vectors produced by running the reference (tests/golden/make_golden_sinc_dvr.py) and the numpy oracle.
Tolerance 1e-12 * max|ref| for the contractions, bit equality for data movement."""

import warnings

import numpy as np
import pytest

from conftest import assert_close_scaled, load_golden
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if complex_ else x


@pytest.fixture(scope="module")
def g():
    return load_golden("sinc_dvr_reference_run")


def make(module, u_repr="2d", l=14, length=6.0):
    from quantum_systems_b200 import ODSincDVR

    return ODSincDVR(l, length, a=0.3, alpha=0.9, beta=0.1, potential=ODSincDVR.DWPotential(1.0, 2.0), u_repr=u_repr,
                     np=module)


def modules():
    from quantum_systems_b200 import xp

    return {"numpy": np, "xp": xp}


@pytest.mark.parametrize("w_complex", [False, True])
@pytest.mark.parametrize("c_complex", [False, True])
@pytest.mark.parametrize("biorth", [False, True])
@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("n,m", [(6, 6), (9, 5), (7, 12), (33, 20)])
def test_diagonal_transform_matches_oracle(n, m, w_complex, c_complex, biorth, anti):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(1000 * n + 10 * m + 4 * w_complex + 2 * c_complex + biorth)
    w = rand(rng, (n, n), w_complex)
    C = rand(rng, (n, m), c_complex)
    Ct = rand(rng, (m, n), c_complex) if biorth else None
    expected = oracle.sinc_dvr_transform_two_body_elements(w, C, Ct, anti_symmetrize=anti)
    got = host(ops.transform_two_body_diagonal(dev(w), dev(C), None if Ct is None else dev(Ct), anti_symmetrize=anti))
    assert got.dtype == expected.dtype
    assert_close_scaled(got, expected, rel=1e-12)


def test_diagonal_transform_equals_dense_transform_at_size():
    """Size-independent property at a size the einsum oracle would not finish quickly: the structured result
    equals the dense four-index transform of the scattered tensor (both on the GPU, independent kernels paths)."""
    from quantum_systems_b200 import ops

    n = 64
    rng = np.random.default_rng(64)
    w = dev(rng.standard_normal((n, n)))
    C = dev(np.linalg.qr(rng.standard_normal((n, n)))[0])
    dense = torch.zeros((n,) * 4, dtype=torch.float64, device="cuda")
    ind = torch.arange(n, device="cuda")
    dense[ind[:, None], ind[None, :], ind[:, None], ind[None, :]] = w
    expected = ops.transform_two_body(dense, C)
    got = ops.transform_two_body_diagonal(w, C)
    assert float((got - expected).abs().max()) <= 1e-12 * float(expected.abs().max())
    anti = ops.transform_two_body_diagonal(w, C, anti_symmetrize=True)
    assert torch.equal(anti, got - got.permute(0, 1, 3, 2))


@pytest.mark.parametrize("module_name", ["numpy", "xp"])
@pytest.mark.parametrize("u_repr", ["2d", "4d"])
def test_setup_matches_reference(g, module_name, u_repr):
    dvr = make(modules()[module_name], u_repr)
    assert dvr.u_repr == u_repr and dvr.sparse_repr == (u_repr == "2d")
    for key in ("h", "s", "u", "spf", "position"):
        got = host(getattr(dvr, key))
        assert got.dtype == g[f"{u_repr}_{key}"].dtype
        np.testing.assert_allclose(got, g[f"{u_repr}_{key}"], atol=1e-14, rtol=0)
    np.testing.assert_array_equal(host(dvr.converted_u("4d")), g["4d_u"])
    np.testing.assert_array_equal(host(dvr.converted_u("2d")), g["2d_u"])
    with pytest.raises(ValueError):
        make(np, "sparse")


@pytest.mark.parametrize("module_name", ["numpy", "xp"])
def test_transform_two_body_elements_override(g, module_name):
    module = modules()[module_name]
    dvr = make(module)
    C, Ct, Cr = (module.asarray(g[k]) for k in ("C", "Ct", "Cr"))
    assert_close_scaled(host(dvr.transform_two_body_elements(dvr.u, C, module)), g["tb_default"])
    assert_close_scaled(host(dvr.transform_two_body_elements(dvr.u, C, module, C_tilde=Ct)), g["tb_biorth"])
    assert_close_scaled(
        host(dvr.transform_two_body_elements(dvr.u, C, module, anti_symmetrize=True, C_tilde=Ct)), g["tb_antisym"]
    )
    assert_close_scaled(host(dvr.transform_two_body_elements(dvr.u, Cr, module)), g["tb_real_C"])
    dense = make(module, "4d")
    assert_close_scaled(host(dense.transform_two_body_elements(dense.u, C, module, C_tilde=Ct)), g["tb_dense_4d"])
    with pytest.raises(AssertionError):
        dense.transform_two_body_elements(dense.u, C, module, anti_symmetrize=True)


@pytest.mark.parametrize("module_name", ["numpy", "xp"])
def test_change_basis_of_the_2d_system(g, module_name):
    module = modules()[module_name]
    dvr = make(module)
    dvr.change_basis(module.asarray(g["C"]), module.asarray(g["Ct"]))
    assert dvr.l == 9 and dvr.u_repr == "4d"
    for key in ("h", "s", "u", "spf", "position"):
        assert_close_scaled(host(getattr(dvr, key)), g[f"cb_{key}"])


@pytest.mark.parametrize("u_repr", ["2d", "4d"])
def test_spin_doubling_uses_the_overridden_hooks(g, u_repr):
    from quantum_systems_b200 import ODSincDVR

    dvr = ODSincDVR(8, 4.0, u_repr=u_repr)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert dvr.change_to_general_orbital_basis(anti_symmetrize=True) is dvr
    assert dvr.l == 16 and dvr.includes_spin and dvr.u_repr == u_repr  # the reference never converts 2d -> 4d
    for key in ("h", "u", "position"):
        np.testing.assert_array_equal(host(getattr(dvr, key)), g[f"spin_{u_repr}_{key}"])
    if u_repr == "2d":
        assert dvr.spin_2_tb is None  # no spin operators for the 2-D storage (reference basis_set.py:578-603)


def test_change_module_is_refused_for_the_2d_storage():
    from quantum_systems_b200 import xp

    dvr = make(np)
    with pytest.warns(UserWarning, match="not implemented for sparse u"):
        dvr.change_module(xp)
    assert dvr.np is xp and isinstance(dvr.u, np.ndarray)
    dense = make(np, "4d")
    dense.change_module(xp)
    assert isinstance(dense.u, torch.Tensor) and dense.u.is_cuda
