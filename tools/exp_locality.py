"""Experiment: GINE aggregation fwd/bwd time on the config-4 graph under different node orderings."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib, graph as G
from raincast_gnn_b200.utils import synthetic as syn
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import reverse_cuthill_mckee

dev = torch.device("cuda:0")
n, h = 100_000, 128
coords = syn.station_coords(n, 1000.0, 0)
ei, ea = G.radius_graph_from_coords(coords, syn.scaled_graph_radius(n, 1000.0))
e = ei.shape[1]
L = _lib.lib()

def morton(c, bits=10):
    q = np.clip((c / 1000.0 * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    def spread(v):
        v = (v | (v << 16)) & 0x0000FFFF0000FFFF
        v = (v | (v << 8)) & 0x00FF00FF00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
        v = (v | (v << 2)) & 0x3333333333333333
        v = (v | (v << 1)) & 0x5555555555555555
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1)

def relabel(order):       # order[new] = old
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    return torch.from_numpy(inv[ei.numpy()])

def bench(name, ei2):
    sg = G.build_station_graph(ei2, ea, n).to(dev)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, generator=g).to(dev); gout = torch.randn(n, h, generator=g).to(dev)
    w, b, eps = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev), torch.zeros(1, device=dev)
    out = torch.empty_like(x)
    nb = L.rc_gine_aggr_bwd_nblocks(n, h); part = torch.empty(nb, 3, h, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def fwd():
        _lib.check(L.rc_gine_aggr_fwd(x.data_ptr(), sg.rowptr.data_ptr(), sg.col.data_ptr(), sg.attr.data_ptr(), w.data_ptr(), b.data_ptr(), eps.data_ptr(), out.data_ptr(), n, h, st))
    def bwd():
        _lib.check(L.rc_gine_aggr_bwd(gout.data_ptr(), x.data_ptr(), sg.t_rowptr.data_ptr(), sg.t_dst.data_ptr(), sg.t_attr.data_ptr(), w.data_ptr(), b.data_ptr(), eps.data_ptr(), None, out.data_ptr(), part.data_ptr(), n, h, st))
    res = []
    for fn in (fwd, bwd):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); c.record(); c.synchronize(); ts.append(a.elapsed_time(c) * 1e3)
        res.append(sum(ts) / len(ts))
    print(f"{name:12s} fwd {res[0]:7.1f} us  bwd {res[1]:7.1f} us   ({126.63e6/res[0]/1e3:6.0f} / {177.83e6/res[1]/1e3:6.0f} GB/s)", flush=True)

bench("original", ei)
bench("morton", relabel(np.argsort(morton(coords), kind="stable")))
t = time.time()
A = coo_matrix((np.ones(e, np.int8), (ei[0].numpy(), ei[1].numpy())), shape=(n, n)).tocsr()
rcm = reverse_cuthill_mckee(A, symmetric_mode=True)
print("rcm time", time.time() - t)
bench("rcm", relabel(np.asarray(rcm, dtype=np.int64)))
