"""A few launches of the links + mixture-CRPS kernel at M = 2^24 (ncu target)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import kernels as K
from raincast_gnn_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
mm = 1 << 24
raw = torch.randn(mm, 5, device=dev)
y = syn.log_precip_targets(mm, seed=5).to(dev)
for _ in range(3):
    K.crps_fwd_bwd(raw, y, 3, raw_input=True)
torch.cuda.synchronize()
print("ok")
