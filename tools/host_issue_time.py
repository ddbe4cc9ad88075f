"""How long does the HOST need to issue one change_basis (no synchronisation inside the loop) next to the device time
of the same calls?  If the host time per call approaches the device time, the launch queue runs dry and the step is
host-bound (development aid)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quantum_systems_b200 import BasisSet, ops, xp


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rng = np.random.default_rng(2)
    u = rng.standard_normal((n,) * 4)
    u = 0.5 * (u + u.transpose(1, 0, 3, 2))
    C = xp.asarray(np.linalg.qr(rng.standard_normal((n, n)))[0])
    for exploit in (False, True):
        ops.EXPLOIT_SYMMETRY = exploit
        basis = BasisSet(n, 1, np=xp)
        basis.h, basis.s, basis.u = rng.standard_normal((n, n)), np.eye(n), u
        for _ in range(3):
            basis.change_basis(C)
        torch.cuda.synchronize()
        steps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            basis.change_basis(C)
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / steps
        torch.cuda.synchronize()
        # the two-body transform alone, and the pieces around it
        t0 = time.perf_counter()
        for _ in range(steps):
            ops.transform_two_body(basis.u, C)
        two_body_host_ms = (time.perf_counter() - t0) * 1e3 / steps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            ops.transform_one_body(basis.h, C)
        one_body_host_ms = (time.perf_counter() - t0) * 1e3 / steps
        torch.cuda.synchronize()
        print(json.dumps({"n": n, "exploit_symmetry": exploit, "device_ms_per_change_basis": e0.elapsed_time(e1) / steps,
                          "host_issue_ms_per_change_basis": host_ms, "host_issue_ms_two_body": two_body_host_ms,
                          "host_issue_ms_one_body": one_body_host_ms}))
    ops.EXPLOIT_SYMMETRY = True


if __name__ == "__main__":
    main()
