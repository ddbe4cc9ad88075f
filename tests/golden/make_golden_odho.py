"""Golden vectors for the harmonic-oscillator grid basis ``ODHO`` (numba trapezoid path), produced by the UNMODIFIED
reference (quantum_systems/quantum_dots/one_dim/one_dim_qd.py:71-166).  The class is not exported by the reference
package and has no test there, so these reference-run vectors are the pin.  Build container
only:

    NUMBA_CACHE_DIR=/tmp/numba PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_odho.py
"""

import os
import sys

import numpy as np

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REFERENCE)
    from quantum_systems.quantum_dots.one_dim.one_dim_qd import ODHO

    out = {}
    cases = {"a": dict(l=6, grid_length=8, num_grid_points=81, omega=0.5, a=0.3, alpha=0.8),
             "b": dict(l=12, grid_length=11, num_grid_points=201, omega=1.0, a=0.25, alpha=1.0),
             "c": dict(l=9, grid_length=10.5, num_grid_points=120, omega=0.25, a=0.25, alpha=1.0)}  # even point count
    for tag, kw in cases.items():
        od = ODHO(**kw)
        for key in ("h", "s", "u", "spf", "position", "grid", "eigen_energies"):
            out[f"{tag}_{key}"] = np.ascontiguousarray(getattr(od, key))
        for key, value in kw.items():
            out[f"{tag}_arg_{key}"] = np.asarray(value)
    np.savez_compressed(os.path.join(HERE, "odho_reference_run.npz"), **out)
    print("wrote odho_reference_run.npz:", {k: v.shape for k, v in out.items() if k.endswith("_u")})


if __name__ == "__main__":
    main()
