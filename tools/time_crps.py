"""Links + mixture-CRPS kernel (value + gradient) at M = 2^24 nodes: CUDA-event time and achieved HBM bandwidth
((2C+1)*4 bytes per node).  python tools/time_crps.py [log2 M]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import kernels as K
from raincast_gnn_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
mm = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
y = syn.log_precip_targets(mm, seed=5).to(dev)
for kind, width in ((3, 5), (2, 4), (1, 3), (0, 2)):
    raw = torch.randn(mm, width, device=dev)
    for _ in range(3):
        K.crps_fwd_bwd(raw, y, kind, raw_input=True)
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        K.crps_fwd_bwd(raw, y, kind, raw_input=True)
    c.record(); c.synchronize()
    t = a.elapsed_time(c) / 5
    by = (2 * width + 1) * 4 * mm
    print(f"crps kind {kind} (C={width}) M=2^{mm.bit_length()-1}: {t*1e3:.1f} us  {by / t / 1e6:.0f} GB/s = {by / t / 1e6 / 6550.7 * 100:.1f} % of 6.55 TB/s")
    del raw
