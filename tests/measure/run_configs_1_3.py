"""BASELINE.json configs[0] and configs[2] end to end through the reference-facing API, next to the numpy oracle
(the reference's call sequence) on the host cores of the same box.  One JSON line per config.

configs[0]: ODQD l = 20, 201 grid points -> GeneralOrbitalSystem (40 spin-orbitals) -> change_basis(random orthonormal C)
configs[2]: ODQD shielded-Coulomb double well l = 100, 2001 grid points: grid interaction build + Fock construction
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import qs_oracle as oracle  # noqa: E402  (CPU baseline / checker)
from quantum_systems_b200 import ODQD, GeneralOrbitalSystem, SpatialOrbitalSystem  # noqa: E402


def wall(fn, reps=3):
    best, out = 1e30, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, out


def host(a):
    return a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def rel(got, ref):
    return float(np.abs(host(got) - ref).max() / np.abs(ref).max())


def config1():
    pot = ODQD.HOPotential(0.25)
    C = np.linalg.qr(np.random.default_rng(1).standard_normal((40, 40)))[0]
    rec = {"config": "configs[0]: ODQD(20, 10, 201, HO 0.25) -> GeneralOrbitalSystem(2) -> change_basis(C 40x40 orthonormal)"}
    for mode in ("xp", "numpy"):
        kw = {} if mode == "xp" else {"np": np}

        def pipeline():
            od = ODQD(20, 10, 201, potential=pot, **kw)
            gos = GeneralOrbitalSystem(2, od)
            gos.change_basis(C if mode == "numpy" else gos.np.asarray(C))
            return gos

        pipeline()  # warm-up (library load, allocator)
        sec, gos = wall(pipeline)
        rec[f"gpu_{mode}_s"] = round(sec, 5)
        rec[f"gpu_{mode}_result"] = gos

    def cpu_pipeline():
        b = oracle.odqd_setup_basis(20, 10, 201, pot)
        g = oracle.change_to_general_orbital_basis({k: b[k] for k in ("h", "s", "u", "position", "spf")})
        return oracle.change_basis(g, C)

    cpu_pipeline()
    t = time.perf_counter()
    ref = cpu_pipeline()
    rec["cpu_oracle_s"] = round(time.perf_counter() - t, 4)
    rec["cpu_cores"] = os.cpu_count()
    for mode in ("xp", "numpy"):
        gos = rec.pop(f"gpu_{mode}_result")
        rec[f"max_rel_err_u_{mode}"] = rel(gos.u, ref["u"])
        rec[f"max_rel_err_h_{mode}"] = rel(gos.h, ref["h"])
        rec[f"speedup_{mode}"] = round(rec["cpu_oracle_s"] / rec[f"gpu_{mode}_s"], 1)
    return rec


def config3():
    pot = ODQD.DWPotential(1.0, 5.0)
    rec = {"config": "configs[2]: ODQD(100, 20, 2001, DW(1, 5)) grid Coulomb build + Fock (spatial n_occ = 10, general n_occ = 20)"}
    ODQD(8, 20, 201, potential=pot)
    sec, od = wall(lambda: ODQD(100, 20, 2001, potential=pot), reps=2)
    rec["gpu_build_s"] = round(sec, 4)
    t = time.perf_counter()
    ref = oracle.odqd_setup_basis(100, 20, 2001, pot)
    rec["cpu_build_s"] = round(time.perf_counter() - t, 3)
    rec["max_rel_err_u"] = rel(od.u, np.ascontiguousarray(ref["u"]))
    spas = SpatialOrbitalSystem(20, od)
    sec, f = wall(lambda: spas.construct_fock_matrix(spas.h, spas.u))
    rec["gpu_fock_spatial_s"] = round(sec, 6)
    t = time.perf_counter()
    f_ref = oracle.construct_fock_matrix_spatial(ref["h"], ref["u"], 10)
    rec["cpu_fock_spatial_s"] = round(time.perf_counter() - t, 6)
    rec["max_rel_err_fock_spatial"] = rel(f, f_ref)
    rec["e_ref_spatial_rel_err"] = abs(spas.compute_reference_energy() - oracle.reference_energy_spatial(ref["h"], ref["u"], 10)) / abs(
        oracle.reference_energy_spatial(ref["h"], ref["u"], 10))
    del spas, od
    rec["cpu_cores"] = os.cpu_count()
    rec["speedup_build"] = round(rec["cpu_build_s"] / rec["gpu_build_s"], 1)
    return rec


if __name__ == "__main__":
    print(json.dumps(config1()), flush=True)
    print(json.dumps(config3()), flush=True)
