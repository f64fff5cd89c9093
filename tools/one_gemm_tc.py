"""One tensor-core GEMM of the config-4 shape, a few launches (ncu target).  python tools/one_gemm_tc.py [fwd|wgrad]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
m, h = 100000, (int(sys.argv[2]) if len(sys.argv) > 2 else 128)
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(m, h, generator=g, device=dev)
dy = torch.randn(m, h, generator=g, device=dev)
w = torch.randn(h, h, generator=g, device=dev) / h ** 0.5
b = torch.randn(h, generator=g, device=dev)
y = torch.empty(m, h, device=dev)
dw, db = torch.empty(h, h, device=dev), torch.empty(h, device=dev)
for _ in range(4):
    if what == "fwd":
        K.gemm(m, h, h, K.operand(x, h), K.operand(w, h), y, h, bias=b, epi=K.RC_EPI_RELU)
    else:
        sink = K.GradSink(dev)
        K.linear_bwd_weight(K.operand(dy, h), K.operand(x, h), m, h, h, dw, db, sink)
        sink.flush()
torch.cuda.synchronize()
print("ok")
