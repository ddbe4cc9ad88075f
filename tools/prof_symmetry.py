"""Driver for the ncu capture of the symmetry-aware path: one exchange-symmetric and one anti-symmetric transform
at n = 96 (detection kernels, masked quarter GEMMs, mirror fill) and the cyclic fill of the sharded variant."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quantum_systems_b200 import _native, ops

n = 96
torch.manual_seed(0)
base = torch.randn((n,) * 4, dtype=torch.float64, device="cuda")
C = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device="cuda"))[0].contiguous()
sym = (0.5 * (base + base.permute(1, 0, 3, 2))).contiguous()
anti = (base - base.permute(0, 1, 3, 2)).contiguous()
out_s = ops.transform_two_body(sym, C)
out_a = ops.transform_two_body(anti, C)
_native.call("qs_cyclic_antisymmetric_fill", ctypes.c_void_p(out_a.data_ptr()), 0, n, n,
             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok", ops.two_body_symmetry(out_s), ops.two_body_symmetry(out_a))
