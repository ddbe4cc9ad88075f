"""Driver for the round-2 `ncu --set full` capture at n = 128 (one GPU): the four full quarter GEMMs, the
symmetry-aware path with 16-byte paired stores and the pipelined mirror fill (exchange-symmetric and anti-symmetric
input), the small-matrix one-body kernel, the pipelined cyclic fill, and scattering launches of an emulated 2-rank
transform (cyclic destinations, source-major T2) with the staged cp.async.bulk epilogue."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quantum_systems_b200 import _native, ops, sharded

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
base = torch.randn((n,) * 4, dtype=torch.float64, device="cuda")
C = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device="cuda"))[0].contiguous()
sym = (0.5 * (base + base.permute(1, 0, 3, 2))).contiguous()
anti = (base - base.permute(0, 1, 3, 2)).contiguous()
ops.transform_two_body(base, C, symmetry=0)           # four full quarter GEMMs
out_s = ops.transform_two_body(sym, C)                # exchange symmetry: detection, masked steps, mirror fill (mode 2)
out_a = ops.transform_two_body(anti, C)               # anti-symmetry: ..., mirror fill (mode 1)
ops.transform_one_body(torch.randn((n, n), dtype=torch.float64, device="cuda"), C)
_native.call("qs_cyclic_antisymmetric_fill", ctypes.c_void_p(out_a.data_ptr()), 0, n, n,
             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
del out_s, base, sym
torch.cuda.empty_cache()
m = 96                                                 # emulated 2-rank sharded transform (scatter to local memory)
ctx = sharded.EmulatedContext(2)
u = torch.randn((m,) * 4, dtype=torch.float64, device="cuda")
basis = sharded.ShardedBasisSet.from_global(ctx, np.eye(m), np.eye(m), u)
Cm = torch.linalg.qr(torch.randn((m, m), dtype=torch.float64, device="cuda"))[0].contiguous()
sharded.transform_two_body_sharded(basis.u, Cm, symmetry=0)
torch.cuda.synchronize()
print("ok")
