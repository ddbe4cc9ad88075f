"""Tiny driver for ncu: the memory-bound passes (fused add_spin + anti-symmetrise, stand-alone anti-symmetrise)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quantum_systems_b200 import ops

l = int(sys.argv[1]) if len(sys.argv) > 1 else 64
u = torch.randn((l,) * 4, dtype=torch.float64, device="cuda")
for _ in range(2):
    a = ops.add_spin_two_body(u, anti_symmetrize=True, out_dtype=torch.float64)
    b = ops.add_spin_two_body(u, anti_symmetrize=True, out_dtype=torch.complex128)
    c = ops.anti_symmetrize(a)
torch.cuda.synchronize()
print("ok", float(c.abs().max()))
