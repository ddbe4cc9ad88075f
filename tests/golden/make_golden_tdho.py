"""Golden vectors for the two-dimensional harmonic-oscillator family, produced by the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is mounted read-only there):

    NUMBA_CACHE_DIR=/tmp/numba PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_tdho.py

Writes
* ``tdho_reference_run.npz``     -- ``quantum_systems`` itself: Coulomb elements l = 12 (full) and the non-zero
  elements of l = 30 as (index, value) lists, one-body / position / spf of a small oscillator, double-well
  and smooth-double-well one-body matrices, the magnetic-field system (levels, h, u, position);
* ``tdho_reference_table.npz``   -- the reference's own golden table
  ``tests/dat/two_dim_quantum_dots_coulomb_elements.dat`` (pinned by tests/test_two_dim_ho.py:70-90),
  ``index_map.dat`` and ``two_dim_quantum_dots_one_body_elements.dat``, re-packed;
* ``tdho_reference_dat.npz``     -- ``tests/dat/{tddw,tdhob}_{h,u,dipole_moment}.npy`` (pinned by
  tests/test_two_dim_dw.py:189-215 and tests/test_two_dim_ho_b_field.py:27-41) plus the eigenvector matrix the
  double-well test derives, so the comparison needs no eigen-solver phase convention.
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


def _sparse(u):
    idx = np.argwhere(u != 0)
    return idx.astype(np.uint8), u[tuple(idx.T)]


def main():
    sys.path.insert(0, REFERENCE)
    import quantum_systems as qs
    from quantum_systems.quantum_dots.two_dim import two_dim_helper as helper

    dat = os.path.join(REFERENCE, "tests", "dat")
    out = {}

    # -- A. Coulomb elements straight from the numba kernel ------------------------------------------
    out["u_l12"] = helper._get_coulomb_elements(12)
    out["u_l30_index"], out["u_l30_value"] = _sparse(helper._get_coulomb_elements(30))
    out["indices_nm"] = np.array([helper.get_indices_nm(p) for p in range(120)], dtype=np.int64)

    # -- B. a small oscillator system: every array setup_basis stores --------------------------------
    ho = qs.TwoDimensionalHarmonicOscillator(10, 4.0, 21, omega=0.7, mass=1.3)
    out.update(ho_h=ho.h, ho_u=ho.u, ho_s=ho.s, ho_position=ho.position, ho_spf=ho.spf)
    gos = qs.GeneralOrbitalSystem(2, ho)
    out.update(ho_gos_u=gos.u, ho_gos_h=gos.h, ho_gos_position=gos.position)

    # -- C. double-well one-body matrices (sympy radial integrals) -----------------------------------
    out["dw_h_axis0"] = helper.get_double_well_one_body_elements(12, 0.8, 1, 3, dtype=np.complex128, axis=0)
    out["dw_h_axis1"] = helper.get_double_well_one_body_elements(12, 1.0, 1, 2, dtype=np.complex128, axis=1)
    out["smooth_dw_h"] = helper.get_smooth_double_well_one_body_elements(8, 0.9, 1, a=2, b=2, dtype=np.complex128)

    # -- D. magnetic field: level table and the spatial system ---------------------------------------
    hob = qs.TwoDimHarmonicOscB(10, 5, 21, omega_c=0.5)
    out["hob_levels"] = hob.df[["n", "m"]].to_numpy().astype(np.int64)
    out["hob_energy"] = hob.df["E"].to_numpy()
    out.update(hob_h=hob.h, hob_u=hob.u, hob_position=hob.position, hob_spf=hob.spf)
    hob2 = qs.TwoDimHarmonicOscB(7, 5, 11, omega=0.6, omega_c=1.3)
    out["hob2_levels"] = hob2.df[["n", "m"]].to_numpy().astype(np.int64)
    out.update(hob2_h=hob2.h, hob2_u=hob2.u)
    _save("tdho_reference_run", **out)

    # -- E. the reference's golden table -------------------------------------------------------------
    rows = np.loadtxt(os.path.join(dat, "two_dim_quantum_dots_coulomb_elements.dat"))
    index_map = np.loadtxt(os.path.join(dat, "index_map.dat"), skiprows=1, dtype=np.int64)
    one_body = np.loadtxt(os.path.join(dat, "two_dim_quantum_dots_one_body_elements.dat"), skiprows=1)
    _save(
        "tdho_reference_table",
        coulomb_index=rows[:, :4].astype(np.uint8),
        coulomb_value=rows[:, 4],
        index_map=index_map,
        one_body_index=one_body[:, 0].astype(np.int64),
        one_body_value=one_body[:, 1],
    )

    # -- F. tests/dat goldens of the double-well and magnetic-field systems --------------------------
    h_dw = helper.get_double_well_one_body_elements(10, 0.8, 1, 3, dtype=np.complex128, axis=0)
    _, C_dw = np.linalg.eigh(h_dw)
    packed = {"tddw_C": qs.BasisSet.add_spin_one_body(C_dw, np=np)}
    for name in ("tddw_h", "tddw_u", "tddw_dipole_moment", "tdhob_h", "tdhob_u", "tdhob_dipole_moment"):
        packed[name] = np.load(os.path.join(dat, name + ".npy"))
    _save("tdho_reference_dat", **packed)


if __name__ == "__main__":
    main()
