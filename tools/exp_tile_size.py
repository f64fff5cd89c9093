"""Experiment: station-tile size (staged rows per tile) against the fixed cost of a config-4 launch.  Smaller tiles
shorten the first-tile ramp and the last-tile imbalance (1 440 tiles over 148 CTAs at the shipped size) and stage more
halo rows per owned row.  L2 flushed between launches; CUDA events; median.  usage: exp_tile_size.py [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import graph as G, kernels as K
from raincast_gnn_b200.utils import synthetic as syn

dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
h, n = 128, 100_000
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); c.synchronize(); ts.append(a.elapsed_time(c) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
sg = G.build_station_graph(ei, ea, n).to(dev)
e = ei.shape[1]
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, generator=g).to(dev); gout = torch.randn(n, h, generator=g).to(dev)
w, b, eps = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev), torch.zeros(1, device=dev)
out = torch.empty_like(x)
ref = torch.empty_like(x)
lim = G.tile_limits(h)
print(f"limits: max_src {lim[0]}, max_block_bytes {lim[1]}", flush=True)
K.gine_aggr_fwd(x, sg, w, b, eps, ref, tiled=False)
for max_src in (lim[0], 150, 128, 104, 80, 60):
    fwd = G.build_tiles_host(sg.rowptr, sg.col, sg.attr, max_src, lim[1], G.TILE_ROW_BYTES).to(dev)
    bwd = G.build_tiles_host(sg.t_rowptr, sg.t_dst, sg.t_attr, max_src, lim[1], G.TILE_ROW_BYTES).to(dev)
    sg.__dict__["_tiles"] = {"pair": (fwd, bwd)}
    tf = timeit(lambda: K.gine_aggr_fwd(x, sg, w, b, eps, out, tiled=True))
    err = float((out - ref).abs().max() / ref.abs().max())
    tb = timeit(lambda: K.gine_aggr_bwd(gout, x, sg, w, b, eps, None, out, tiled=True))
    print(f"max_src {max_src:4d}: {fwd.n_tiles:5d} tiles, {(fwd.n_halo + n) / n:.2f} staged rows/row, max staged {fwd.max_staged}, "
          f"{e / max(fwd.n_entries, 1):.2f} edges/entry | fwd {tf[0]:6.1f} us (min {tf[1]:6.1f}) rel err vs untiled {err:.1e} | "
          f"bwd {tb[0]:6.1f} us (min {tb[1]:6.1f})", flush=True)
