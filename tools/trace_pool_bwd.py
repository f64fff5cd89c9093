"""Per-role timeline of CTA 0 of the tensor-core DeepSets pool backward (debug trace): python tools/trace_pool_bwd.py [members]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
m, em, f, h = 100_000, int(sys.argv[1]) if len(sys.argv) > 1 else 51, 35, 128
g = torch.Generator(device=dev).manual_seed(0)
ens = torch.randn(m, em, f, generator=g, device=dev)
w1 = torch.randn(h, f, generator=g, device=dev) * 0.2
b1 = torch.randn(h, generator=g, device=dev)
dp = torch.randn(m, h, generator=g, device=dev)
nb = int(L.rc_deepsets_pool_bwd_nblocks(m, em, f, h))
part = torch.empty(nb, h * f + h, device=dev)
st = torch.cuda.current_stream().cuda_stream
trace = torch.zeros(8, 32, dtype=torch.int64, device=dev)


def run():
    _lib.check(L.rc_deepsets_pool_bwd(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), dp.data_ptr(), part.data_ptr(), m, em, f, h, 0, None, st))


run()
torch.cuda.synchronize()
L.rc_debug_ds_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
L.rc_debug_ds_trace(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
names = ["convert begin", "convert end", "MMA1 issue", "MMA2 issue", "E1 begin", "E1 end", "E1 after flush"]
for ev, name in enumerate(names):
    print(f"{name:16s}", " ".join(f"{int(t[ev, i]) - t0:6d}" for i in range(10, 22)))
