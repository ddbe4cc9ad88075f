"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is mounted read-only there and does not exist
on the GPU box):

    NUMBA_CACHE_DIR=/tmp/numba PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every array written here comes out of ``quantum_systems`` as imported from ``/root/reference``;
inputs are stored next to outputs so the tests never need the reference (or a particular numpy
random stream) again.  The last block re-packs the reference's own golden files
(``tests/dat/od*_{h,u,spf,dipole_moment}.npy``, pinned by ``tests/test_one_dim_qd.py:127-143``) as
compressed ``.npz`` so they travel with the repo.
"""

import os
import sys

import numpy as np

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz: {os.path.getsize(path) / 1024:.1f} KiB, keys={sorted(arrays)}")


def _basis_arrays(bs, prefix):
    out = {}
    for key in ("h", "s", "u", "position", "spf", "spin_x", "spin_y", "spin_z", "spin_2", "spin_2_tb"):
        val = getattr(bs, key, None)
        if val is not None:
            out[prefix + key] = np.ascontiguousarray(val)
    return out


def main():
    sys.path.insert(0, REFERENCE)
    import quantum_systems as qs
    from quantum_systems import BasisSet, GeneralOrbitalSystem, ODQD, RandomBasisSet, SpatialOrbitalSystem

    # -- A. bare four-index / one-body transforms (tests/test_helper.py:14-69) -------------------
    rng = np.random.default_rng(1001)
    l = 10
    u = rng.random((l, l, l, l)) + 1j * rng.random((l, l, l, l))
    h = rng.random((l, l)) + 1j * rng.random((l, l))
    C = rng.random((l, l)) + 1j * rng.random((l, l))
    _save(
        "transform_square_complex",
        u=u,
        h=h,
        C=C,
        u_out=BasisSet.transform_two_body_elements(u, C, np=np),
        h_out=BasisSet.transform_one_body_elements(h, C, np=np),
    )

    # real, rectangular 9 -> 6 (odd sizes on purpose)
    rng = np.random.default_rng(1002)
    u = rng.standard_normal((9, 9, 9, 9))
    h = rng.standard_normal((9, 9))
    C = rng.standard_normal((9, 6))
    _save(
        "transform_rect_real",
        u=u,
        h=h,
        C=C,
        u_out=BasisSet.transform_two_body_elements(u, C, np=np),
        h_out=BasisSet.transform_one_body_elements(h, C, np=np),
    )

    # bi-orthogonal: C_tilde = C^-1 != C^dagger (never exercised by the reference's tests)
    rng = np.random.default_rng(1003)
    l = 8
    u = rng.standard_normal((l, l, l, l)) + 1j * rng.standard_normal((l, l, l, l))
    h = rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l))
    q1, _ = np.linalg.qr(rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l)))
    q2, _ = np.linalg.qr(rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l)))
    C = q1 @ np.diag(rng.uniform(0.5, 2.0, l)) @ q2
    C_tilde = np.linalg.inv(C)
    _save(
        "transform_biorthogonal",
        u=u,
        h=h,
        C=C,
        C_tilde=C_tilde,
        u_out=BasisSet.transform_two_body_elements(u, C, np=np, C_tilde=C_tilde),
        h_out=BasisSet.transform_one_body_elements(h, C, np=np, C_tilde=C_tilde),
    )

    # real u, complex C (mixed dtype promotion)
    rng = np.random.default_rng(1004)
    l = 8
    u = rng.standard_normal((l, l, l, l))
    C = rng.standard_normal((l, 6)) + 1j * rng.standard_normal((l, 6))
    _save("transform_real_u_complex_C", u=u, C=C, u_out=BasisSet.transform_two_body_elements(u, C, np=np))

    # -- B. add_spin / anti-symmetrise (tests/test_helper.py:72-147) -----------------------------
    rng = np.random.default_rng(1005)
    l = 5
    u = rng.random((l, l, l, l))
    u = u + u.transpose(1, 0, 3, 2)
    h = rng.random((l, l))
    u_spin = BasisSet.add_spin_two_body(u, np=np)
    _save(
        "add_spin_antisym_real",
        u=u,
        h=h,
        h_spin=BasisSet.add_spin_one_body(h, np=np),
        u_spin=u_spin,
        u_as=BasisSet.anti_symmetrize_u(u_spin),
    )
    uc = u + 1j * rng.random((l, l, l, l))
    uc_spin = BasisSet.add_spin_two_body(uc, np=np)
    _save("add_spin_antisym_complex", u=uc, u_spin=uc_spin, u_as=BasisSet.anti_symmetrize_u(uc_spin))

    # -- C. systems: spatial -> general, rectangular change_basis, Fock, E_ref -------------------
    #      (tests/test_custom_system.py:26-68)
    np.random.seed(2001)
    n, l, dim = 2, 5, 2
    new_l = 8
    spas = SpatialOrbitalSystem(n, RandomBasisSet(l, dim))
    arrays = _basis_arrays(spas._basis_set, "spas_")
    arrays["spas_nuclear_repulsion_energy"] = np.asarray(spas.nuclear_repulsion_energy)
    gos = spas.construct_general_orbital_system()
    arrays.update(_basis_arrays(gos._basis_set, "gos_"))
    arrays["spas_fock"] = spas.construct_fock_matrix(spas.h, spas.u)
    arrays["gos_fock"] = gos.construct_fock_matrix(gos.h, gos.u)
    arrays["spas_e_ref"] = np.asarray(spas.compute_reference_energy())
    arrays["gos_e_ref"] = np.asarray(gos.compute_reference_energy())
    C_spas = RandomBasisSet.get_random_elements((spas.l, new_l), np)
    C_gos = RandomBasisSet.get_random_elements((gos.l, new_l), np)
    arrays["C_spas"] = C_spas
    arrays["C_gos"] = C_gos
    spas.change_basis(C_spas)
    gos.change_basis(C_gos)
    arrays.update(_basis_arrays(spas._basis_set, "spas_cb_"))
    arrays.update(_basis_arrays(gos._basis_set, "gos_cb_"))
    arrays["gos_cb_fock"] = gos.construct_fock_matrix(gos.h, gos.u)
    arrays["n"] = np.asarray(n)
    _save("systems_random", **arrays)

    # -- D. ODQD grid build (tests/test_one_dim_qd.py) at a small size ---------------------------
    for tag, pot, kw in [
        ("ho", ODQD.HOPotential(1.0), dict(l=6, grid_length=5, num_grid_points=101)),
        ("dw", ODQD.DWPotential(1.0, 5.0), dict(l=7, grid_length=6, num_grid_points=128, a=0.3, alpha=0.9, beta=0.1)),
    ]:
        od = ODQD(potential=pot, **kw)
        arrays = dict(
            h=od.h,
            s=od.s,
            u=np.ascontiguousarray(od.u),
            spf=od.spf,
            position=od.position,
            eigen_energies=od.eigen_energies,
            grid=od.grid,
        )
        spas = SpatialOrbitalSystem(2, od.copy_basis())
        arrays["spas_fock"] = spas.construct_fock_matrix(spas.h, spas.u)
        arrays["spas_e_ref"] = np.asarray(spas.compute_reference_energy())
        gos = GeneralOrbitalSystem(2, od)
        arrays["gos_u"] = gos.u
        arrays["gos_h"] = gos.h
        arrays["gos_fock"] = gos.construct_fock_matrix(gos.h, gos.u)
        arrays["gos_e_ref"] = np.asarray(gos.compute_reference_energy())
        _save("odqd_small_" + tag, **arrays)

    # config 1 of BASELINE.json at its real size: ODQD(20, 10, 201) -> 40 spin-orbitals -> change_basis
    od = ODQD(20, 10, 201, potential=ODQD.HOPotential(0.25))
    gos = GeneralOrbitalSystem(2, od)
    C, _ = np.linalg.qr(np.random.default_rng(1).standard_normal((40, 40)))
    gos.change_basis(C)
    # keep the fixture small: a strided sample of u plus full h
    idx = np.arange(0, 40, 3)
    _save(
        "config1_odqd40_change_basis",
        C=C,
        h=gos.h,
        u_sample=np.ascontiguousarray(gos.u[np.ix_(idx, idx, idx, idx)]),
        idx=idx,
        u_abs_sum=np.asarray(np.abs(gos.u).sum()),
        u_max=np.asarray(np.abs(gos.u).max()),
    )

    # -- E. the reference's own golden files, re-packed (data, not source) -----------------------
    dat = os.path.join(REFERENCE, "tests", "dat")
    for name in ("odho", "oddw", "odgauss", "oddw_smooth"):
        _save(
            "ref_dat_" + name,
            h=np.load(os.path.join(dat, name + "_h.npy")),
            u=np.load(os.path.join(dat, name + "_u.npy")),
            spf=np.load(os.path.join(dat, name + "_spf.npy")),
            dipole_moment=np.load(os.path.join(dat, name + "_dipole_moment.npy")),
        )
    print("reference version:", getattr(qs, "__version__", "unknown"), "numpy", np.__version__)


if __name__ == "__main__":
    main()
