"""Tiny driver for ncu: a few four-index transforms at size n (real unless 'c' is given)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantum_systems_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dt = torch.complex128 if len(sys.argv) > 2 and sys.argv[2] == "c" else torch.float64
u = torch.randn((n,) * 4, dtype=dt, device="cuda")
C = torch.linalg.qr(torch.randn((n, n), dtype=dt, device="cuda"))[0].contiguous()
for _ in range(2):
    out = ops.transform_two_body(u, C)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
