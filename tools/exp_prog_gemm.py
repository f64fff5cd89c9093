"""Experiment: per-phase cost of a GEMM / aggregation inside the step program vs as standalone kernels."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib, kernels as K, graph as G
import bench as B
dev = torch.device("cuda:0")
L = _lib.lib()
m, h = 976, 128
x = torch.randn(m, h, device=dev); w = torch.randn(h, h, device=dev); b = torch.randn(h, device=dev)
ei, ea, ei_b, ea_b = B.static_graph(8)
sg = G.build_station_graph(ei_b, ea_b, m).to(dev)
we, be, eps = torch.randn(h, device=dev), torch.randn(h, device=dev), torch.zeros(1, device=dev)
out = torch.empty(m, h, device=dev)
st = torch.cuda.current_stream().cuda_stream

def aggr():
    _lib.check(L.rc_gine_aggr_fwd(x.data_ptr(), sg.rowptr.data_ptr(), sg.col.data_ptr(), sg.attr.data_ptr(), we.data_ptr(), be.data_ptr(), eps.data_ptr(), out.data_ptr(), m, h, torch.cuda.current_stream().cuda_stream))

def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); c.synchronize()
    return a.elapsed_time(c) / n * 1e3

N = 50
for name, body in (("gemm 976x128x128", lambda: K.linear_fwd(x, w, b, relu=True)), ("gine aggr fwd", aggr)):
    # standalone kernels in a CUDA graph
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): body()
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(N): body()
    t_graph = timeit(g.replay) / N
    # the same N calls as a step program
    _lib.check(L.rc_prog_begin()); K.RECORD.active, K.RECORD.keep = True, []
    for _ in range(N): body()
    keep = K.RECORD.keep; K.RECORD.active = False
    nb = L.rc_prog_bytes(); buf = torch.empty(nb, dtype=torch.uint8, device=dev); info = (C.c_int * 4)()
    _lib.check(L.rc_prog_end(buf.data_ptr(), nb, info))
    t_prog = timeit(lambda: L.rc_prog_run(buf.data_ptr(), info, st)) / N
    print(f"{name:18s}: graph {t_graph:6.2f} us/op   program {t_prog:6.2f} us/phase  (ops {info[0]} phases {info[1]} smem {info[2]})")
