"""Device time of the 2-D harmonic-oscillator Coulomb kernel (qs_tdho_coulomb) next to the CPU oracle
(C restatement of the reference algorithm, OpenMP on all host cores).  Prints one JSON line per basis size."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tdho  # noqa: E402  (checker / CPU baseline only)
from quantum_systems_b200 import ops  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [36, 66, 105]
    for l in sizes:
        n, m = tdho.quantum_numbers(l)
        ops.tdho_coulomb(n, m)
        torch.cuda.synchronize()
        times = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            u = ops.tdho_coulomb(n, m)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        nonzero = int((u != 0).sum())
        rec = {"l": l, "elements": l**4, "nonzero": nonzero, "gpu_ms_best": min(times), "gpu_ms_all": times,
               "out_gb": 8 * l**4 / 1e9}
        if l <= 36:
            t = time.perf_counter()
            ref = tdho.get_coulomb_elements(l)
            rec["cpu_oracle_s"] = time.perf_counter() - t
            rec["cpu_cores"] = os.cpu_count()
            rec["max_abs_diff_vs_oracle"] = float(np.abs(u.cpu().numpy() - ref).max())
            rec["speedup_vs_cpu_oracle"] = rec["cpu_oracle_s"] * 1e3 / rec["gpu_ms_best"]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
