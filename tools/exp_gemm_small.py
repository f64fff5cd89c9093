"""Experiment: latency of one small GEMM (976 x 128 x 128) as a function of the row tile."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import kernels as K
dev = torch.device("cuda:0")
N = 50
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); c.synchronize()
    return a.elapsed_time(c) / n * 1e3
for (m, n, k) in ((976, 128, 128), (976, 128, 163), (7808, 128, 128)):
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev); y = torch.empty(m, n, device=dev)
    dy = torch.randn(m, n, device=dev)
    for rm in (1, 2, 4, 8):
        res = []
        for name, body in (("fwd", lambda: K.gemm(m, n, k, K.operand(x, k), K.operand(w, k), y, n, bias=b, epi=K.RC_EPI_RELU, rows_per_warp=rm)),
                           ("bwdD", lambda: K.gemm(m, k, n, K.operand(dy, n), K.operand(w, k), x, k, b_layout=K.RC_B_RED, rows_per_warp=rm))):
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(3): body()
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(N): body()
            res.append(timeit(g.replay) / N)
        print(f"m={m} n={n} k={k} rm={rm}: fwd {res[0]:6.2f} us  bwdD {res[1]:6.2f} us")
