"""CUDA-event time of the bf16 tcgen05 DeepSets contraction at the config-5 shape (H = 512) + check against float64."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
m, em, f, h = 100_000, 51, 35, 512
ens = torch.randn(m, em, f, device=dev); w1 = torch.randn(h, f, device=dev) * 0.2; b1 = torch.randn(h, device=dev)
pooled = torch.empty(m, h, device=dev)
run = lambda: _lib.check(L.rc_deepsets_pool_fwd_bf16(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), pooled.data_ptr(), m, em, f, h, st))
for _ in range(2): run()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
c.record(); c.synchronize()
t = a.elapsed_time(c) / 5
n = 4000
ref = torch.relu(ens[:n].bfloat16().double() @ w1.bfloat16().double().T + b1.double()).sum(1)
err = ((pooled[:n].double() - ref).abs().max() / ref.abs().max()).item()
tail = torch.relu(ens[-n:].bfloat16().double() @ w1.bfloat16().double().T + b1.double()).sum(1)
err2 = ((pooled[-n:].double() - tail).abs().max() / tail.abs().max()).item()
print(f"bf16 h=512 contraction: {t*1e3:.1f} us, {2.0*m*em*f*h/t/1e9:.1f} TFLOP/s algorithmic; rel err vs bf16-rounded float64: {err:.2e} / {err2:.2e}")
