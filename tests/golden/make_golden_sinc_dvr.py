"""Golden vectors for the one-dimensional sinc-DVR basis, produced by the UNMODIFIED reference
(``quantum_systems.ODSincDVR``, sinc_dvr/one_dim/sinc_dvr.py).  The reference has no test of this class
(SURVEY.md section 2 row 7), so these reference-run vectors are the pin.  Build container only:

    NUMBA_CACHE_DIR=/tmp/numba PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sinc_dvr.py
"""

import os
import sys
import warnings

import numpy as np

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REFERENCE)
    import quantum_systems as qs

    out = {}
    l, length = 14, 6.0
    for repr_ in ("2d", "4d"):
        dvr = qs.ODSincDVR(l, length, a=0.3, alpha=0.9, beta=0.1, potential=qs.ODSincDVR.DWPotential(1.0, 2.0), u_repr=repr_)
        for key in ("h", "s", "u", "spf", "position"):
            out[f"{repr_}_{key}"] = np.ascontiguousarray(getattr(dvr, key))
    out["grid"] = dvr.grid

    rng = np.random.default_rng(77)
    C = rng.standard_normal((l, 9)) + 1j * rng.standard_normal((l, 9))
    Ct = rng.standard_normal((9, l)) + 1j * rng.standard_normal((9, l))
    out["C"], out["Ct"] = C, Ct
    dvr = qs.ODSincDVR(l, length, a=0.3, alpha=0.9, beta=0.1, potential=qs.ODSincDVR.DWPotential(1.0, 2.0))
    out["tb_default"] = dvr.transform_two_body_elements(dvr.u, C, np)
    out["tb_biorth"] = dvr.transform_two_body_elements(dvr.u, C, np, C_tilde=Ct)
    out["tb_antisym"] = dvr.transform_two_body_elements(dvr.u, C, np, anti_symmetrize=True, C_tilde=Ct)
    Cr = np.linalg.qr(rng.standard_normal((l, l)))[0]
    out["Cr"] = Cr
    out["tb_real_C"] = dvr.transform_two_body_elements(dvr.u, Cr, np)
    # the 2d path agrees with the dense transform of the 4d representation
    dense = qs.ODSincDVR(l, length, a=0.3, alpha=0.9, beta=0.1, potential=qs.ODSincDVR.DWPotential(1.0, 2.0), u_repr="4d")
    out["tb_dense_4d"] = dense.transform_two_body_elements(dense.u, C, np, C_tilde=Ct)
    # whole change_basis of the 2d system
    dvr.change_basis(C, Ct)
    for key in ("h", "s", "u", "spf", "position"):
        out[f"cb_{key}"] = np.ascontiguousarray(getattr(dvr, key))
    # spin doubling, both representations (called on the basis set: GeneralOrbitalSystem passes a=, b= which
    # ODSincDVR.change_to_general_orbital_basis does not accept, sinc_dvr.py:178-189)
    for repr_ in ("2d", "4d"):
        dvr = qs.ODSincDVR(8, 4.0, u_repr=repr_)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dvr.change_to_general_orbital_basis(anti_symmetrize=True)
        for key in ("h", "u", "position"):
            out[f"spin_{repr_}_{key}"] = np.ascontiguousarray(getattr(dvr, key))
    path = os.path.join(HERE, "sinc_dvr_reference_run.npz")
    np.savez_compressed(path, **out)
    print(f"{os.path.basename(path)}: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


if __name__ == "__main__":
    main()
