"""Launches for ncu evidence of the two other hot kernels of the path (north_star): the links + mixture-CRPS kernel at
M = 2^24 nodes, and the tcgen05 DeepSets member contraction at the config-4 shape (3xTF32, fp32-accurate) and at the
config-5 shape (bf16 operands, H = 512).   ncu -k regex:"crps_main|deepsets_tc|ds_tc" ... python tools/prof_misc.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib, kernels as K
from raincast_gnn_b200.utils import synthetic as syn
L = _lib.lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
mm = 1 << 24
raw = torch.randn(mm, 5, device=dev)
y = syn.log_precip_targets(mm, seed=5).to(dev)
for _ in range(3):
    K.crps_fwd_bwd(raw, y, 3, raw_input=True)
del raw, y
g = torch.Generator().manual_seed(0)
for h, bf16 in ((128, False), (512, True)):
    m, em, f = 100_000, 51, 35
    ens = torch.randn(m, em, f, generator=g).to(dev)
    w1 = (torch.randn(h, f, generator=g) * 0.2).to(dev); b1 = torch.randn(h, generator=g).to(dev)
    pooled = torch.empty(m, h, device=dev)
    fn = L.rc_deepsets_pool_fwd_bf16 if bf16 else L.rc_deepsets_pool_fwd
    for _ in range(3):
        _lib.check(fn(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), pooled.data_ptr(), m, em, f, h, st))
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        _lib.check(fn(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), pooled.data_ptr(), m, em, f, h, st))
    c.record(); c.synchronize()
    t = a.elapsed_time(c) / 5
    flops = 2.0 * m * em * f * h
    print(f"deepsets member contraction m={m} em={em} f={f} h={h} {'bf16' if bf16 else '3xTF32'}: {t*1e3:.1f} us, "
          f"{flops / t / 1e9:.1f} TFLOP/s algorithmic, ens read {ens.numel()*4/t/1e6:.0f} GB/s")
    del ens, pooled
