"""DeepSets pool backward (mask recomputation + d W1 / d b1 partials) at the config-4 shape: CUDA-event time, target for ncu.
   ncu -k regex:deepsets_pool_bwd ... python tools/prof_pool_bwd_scale.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
m, em, f, h = 100_000, int(sys.argv[1]) if len(sys.argv) > 1 else 51, 35, 128
g = torch.Generator().manual_seed(0)
ens = torch.randn(m, em, f, generator=g).to(dev)
w1 = (torch.randn(h, f, generator=g) * 0.2).to(dev); b1 = torch.randn(h, generator=g).to(dev)
dp = torch.randn(m, h, generator=g).to(dev)
nb = int(L.rc_deepsets_pool_bwd_nblocks(m, em, f, h))
part = torch.empty(nb, h * f + h, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(L.rc_deepsets_pool_bwd(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), dp.data_ptr(), part.data_ptr(), m, em, f, h, 0, None, st))
for _ in range(2): run()
torch.cuda.synchronize()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
c.record(); c.synchronize()
t = a.elapsed_time(c) / 5
flops = 4.0 * m * em * f * h
print(f"deepsets pool backward m={m} em={em} f={f} h={h}: {t*1e3:.1f} us, nb={nb}, {flops / t / 1e9:.1f} TFLOP/s algorithmic (fp32 FFMA), "
      f"ens read {ens.numel()*4/t/1e6:.0f} GB/s")
# check against a float64 reference on a slice of the rows (partials summed over CTAs)
dw = part[:, :h * f].sum(0).reshape(h, f).double().cpu()
pre = ens.reshape(-1, f).double() @ w1.double().T + b1.double()
dh = dp.double().repeat_interleave(em, 0) * (pre > 0)
want = (dh.T @ ens.reshape(-1, f).double()).cpu()
print("max rel err d W1 vs float64: %.2e" % ((dw - want).abs().max() / want.abs().max()).item())
