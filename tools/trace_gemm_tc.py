"""Per-role timeline of CTA 0 of the tensor-core rows kernel (debug trace): python tools/trace_gemm_tc.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib, kernels as K  # noqa: E402

dev = torch.device("cuda:0")
m, h = 100000, 128
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(m, h, generator=g, device=dev)
w = torch.randn(h, h, generator=g, device=dev) / h ** 0.5
b = torch.randn(h, generator=g, device=dev)
y = torch.empty(m, h, device=dev)
L = _lib.lib()
L.rc_debug_tc_trace.argtypes = [C.c_void_p]
L.rc_debug_tc_trace.restype = None
trace = torch.zeros(3, 16, 2, dtype=torch.int64, device=dev)
for _ in range(3):
    K.gemm(m, h, h, K.operand(x, h), K.operand(w, h), y, h, bias=b, epi=K.RC_EPI_RELU)
torch.cuda.synchronize()
L.rc_debug_tc_trace(trace.data_ptr())
K.gemm(m, h, h, K.operand(x, h), K.operand(w, h), y, h, bias=b, epi=K.RC_EPI_RELU)
torch.cuda.synchronize()
L.rc_debug_tc_trace(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
print("w staging:", [(int(t[2, i, 0]) - t0, int(t[2, i, 1]) - t0) for i in range(8, 12)])
for role, name in enumerate(("producer", "mma", "epilogue")):
    print(name, " ".join(f"[{int(t[role, i, 0]) - t0:6d},{int(t[role, i, 1]) - t0:6d}]" for i in range(8) if int(t[role, i, 0]) > 0))
