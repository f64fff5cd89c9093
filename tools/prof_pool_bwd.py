"""A few launches of the DeepSets pool backward / forward kernels at the reference shape (target for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from raincast_gnn_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
m, em, f, h = 8 * B.N_STATIONS, B.MEMBERS, B.FEATS, B.HIDDEN
g = torch.Generator().manual_seed(0)
ens = torch.randn(m, em, f, generator=g).to(dev); w1 = (torch.randn(h, f, generator=g) * 0.2).to(dev); b1 = torch.randn(h, generator=g).to(dev)
dp = torch.randn(m, h, generator=g).to(dev); pooled = torch.empty(m, h, device=dev)
nb = int(L.rc_deepsets_pool_bwd_nblocks(m, em, f, h))
part = torch.empty(nb, h * f + h, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    _lib.check(L.rc_deepsets_pool_fwd(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), pooled.data_ptr(), m, em, f, h, st))
    _lib.check(L.rc_deepsets_pool_bwd(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), dp.data_ptr(), part.data_ptr(), m, em, f, h, 0, None, st))
torch.cuda.synchronize()
print("nb", nb)
