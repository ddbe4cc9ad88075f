"""Every operator of the hot path on the GPU next to the numpy oracle (the reference's own call sequence) on the
box's host cores, same inputs, same run -- the per-operator table of SURVEY.md section 8d.  One JSON line per
(operator, size): device milliseconds (CUDA events, best of 5 after warm-up), CPU seconds (best of 2), achieved
TFLOP/s or GB/s against the measured roofline, and the max relative deviation between the two results."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import qs_oracle as oracle  # noqa: E402  (CPU baseline / checker)
from quantum_systems_b200 import ops  # noqa: E402


def gpu_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def cpu_s(fn, reps=2):
    best, out = 1e30, None
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best, out


def rel(got, ref):
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [40, 64, 96, 128]
    peak = ops.probe_dmma_tflops()
    hbm = ops.probe_copy_gbs(1 << 30)
    try:
        from threadpoolctl import threadpool_info

        threads = max([p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"] or [os.cpu_count()])
    except Exception:
        threads = os.cpu_count()
    print(json.dumps({"dmma_peak_tflops": peak, "hbm_copy_gbs": hbm, "host_blas_threads": threads,
                      "host_cpus": os.cpu_count()}), flush=True)
    rng = np.random.default_rng(0)
    for n in sizes:
        for cplx in (False, True):
            if cplx and n > 96:
                continue  # the complex oracle at n = 128 needs > 4 x 4.3 GB of host temporaries and ~20 s; skipped
            u = rng.standard_normal((n,) * 4) + (1j * rng.standard_normal((n,) * 4) if cplx else 0)
            C = np.linalg.qr(rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0))[0]
            u_dev, C_dev = torch.from_numpy(u).cuda(), torch.from_numpy(C).cuda()
            ms, out = gpu_ms(lambda: ops.transform_two_body(u_dev, C_dev))
            sec, ref = cpu_s(lambda: oracle.transform_two_body_elements(u, C), reps=1 if n >= 96 else 2)
            flops = 8.0 * n**5 * (4 if cplx else 1)
            print(json.dumps({"op": "transform_two_body_elements", "n": n, "complex": cplx, "gpu_ms": round(ms, 4),
                              "cpu_s": round(sec, 4), "speedup": round(sec * 1e3 / ms, 1),
                              "gpu_tflops": round(flops / ms * 1e-9, 2), "frac_of_dmma_peak": round(flops / ms * 1e-9 / peak, 3),
                              "cpu_tflops": round(flops / sec * 1e-12, 4), "max_rel_dev": rel(out.cpu().numpy(), ref)}), flush=True)
            del u_dev, out, ref
        l = n // 2
        us = rng.standard_normal((l,) * 4)
        us_dev = torch.from_numpy(us).cuda()
        ms, out = gpu_ms(lambda: ops.add_spin_two_body(us_dev, anti_symmetrize=True, out_dtype=torch.complex128))
        sec, ref = cpu_s(lambda: oracle.anti_symmetrize_u(oracle.add_spin_two_body(us)).astype(np.complex128), reps=1)
        nbytes = 8 * l**4 + 16 * n**4
        print(json.dumps({"op": "add_spin + anti_symmetrize_u + cast_to_complex", "l": l, "n": n, "gpu_ms": round(ms, 4),
                          "cpu_s": round(sec, 4), "speedup": round(sec * 1e3 / ms, 1), "gpu_gbs": round(nbytes / ms * 1e-6, 1),
                          "frac_of_hbm_copy_peak": round(nbytes / ms * 1e-6 / hbm, 3),
                          "exact": bool(np.array_equal(out.cpu().numpy(), ref))}), flush=True)
        a_dev = out
        a = ref
        ms, out = gpu_ms(lambda: ops.anti_symmetrize(a_dev))
        sec, ref2 = cpu_s(lambda: oracle.anti_symmetrize_u(a), reps=1)
        print(json.dumps({"op": "anti_symmetrize_u (complex128)", "n": n, "gpu_ms": round(ms, 4), "cpu_s": round(sec, 4),
                          "speedup": round(sec * 1e3 / ms, 1), "gpu_gbs": round(32 * n**4 / ms * 1e-6, 1),
                          "exact": bool(np.array_equal(out.cpu().numpy(), ref2))}), flush=True)
        del ref2, out
        n_occ = max(2, n // 10)
        h = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        h_dev = torch.from_numpy(h).cuda()
        ms, out = gpu_ms(lambda: ops.fock_general(h_dev, a_dev, n_occ))
        sec, ref3 = cpu_s(lambda: oracle.construct_fock_matrix_general(h, a, n_occ))
        print(json.dumps({"op": "construct_fock_matrix (general)", "n": n, "n_occ": n_occ, "gpu_ms": round(ms, 4),
                          "cpu_s": round(sec, 5), "speedup": round(sec * 1e3 / ms, 1),
                          "max_rel_dev": rel(out.cpu().numpy(), ref3)}), flush=True)
        del a_dev, a, us_dev
    for l, G in ((20, 201), (50, 1001), (100, 2001)):
        grid, eps, C = oracle.odqd_orbitals(l, 10.0 if l == 20 else 20.0, G, lambda x: 0.5 * 0.0625 * x**2)
        C_dev, g_dev = torch.from_numpy(C).cuda(), torch.from_numpy(grid[1:-1].copy()).cuda()
        ms, out = gpu_ms(lambda: ops.odqd_coulomb(C_dev, g_dev, 1.0, 0.25), reps=3, warm=1)
        sec, ref = cpu_s(lambda: np.ascontiguousarray(oracle.odqd_coulomb_elements(C, grid, 1.0, 0.25)), reps=1)
        Gp = G - 2
        flops = 2.0 * l**2 * Gp**2 + 2.0 * l**4 * Gp
        print(json.dumps({"op": "ODQD grid Coulomb build", "l": l, "G": G, "gpu_ms": round(ms, 4), "cpu_s": round(sec, 4),
                          "speedup": round(sec * 1e3 / ms, 1), "gpu_tflops": round(flops / ms * 1e-9, 2),
                          "frac_of_dmma_peak": round(flops / ms * 1e-9 / peak, 3),
                          "max_rel_dev": rel(out.cpu().numpy(), ref)}), flush=True)


if __name__ == "__main__":
    main()
