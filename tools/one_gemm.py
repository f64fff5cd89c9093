import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import kernels as K
dev = torch.device("cuda:0")
m, n, k = 976, 128, 128
x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev); y = torch.empty(m, n, device=dev)
for rm in (1, 2, 4):
    for _ in range(20):
        K.gemm(m, n, k, K.operand(x, k), K.operand(w, k), y, n, bias=b, epi=K.RC_EPI_RELU, rows_per_warp=rm)
torch.cuda.synchronize()
