"""Experiment: GINE aggregation, warp-per-row vs station-tile kernels, on the config-4 graph and on a large batch of
reference graphs.  L2 flushed between launches; CUDA events.  usage: exp_tiled.py [iters] [which: c4,c4x10,ref,all]
c4x10: SURVEY.md 8d's optional x10 graph (1 M nodes, same mean degree): x = 512 MB, four times L2, so the gathered rows
come from HBM whatever the flush does."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import _lib, graph as G, kernels as K
from raincast_gnn_b200.utils import synthetic as syn

dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
which = sys.argv[2] if len(sys.argv) > 2 else "all"
h = 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); c.synchronize(); ts.append(a.elapsed_time(c) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def run(name, ei, ea, n):
    e = ei.shape[1]
    t0 = time.time()
    sg = G.build_station_graph(ei, ea, n).to(dev)
    t1 = time.time()
    tiles = sg.tiles(h)
    t2 = time.time()
    print(f"== {name}: {n} nodes, {e} edges; csr build {t1-t0:.2f}s, tiles build {t2-t1:.2f}s: "
          f"fwd {tiles[0].n_tiles} tiles, halo {tiles[0].n_halo} ({(tiles[0].n_halo + n) / n:.2f} staged rows/row), max staged {tiles[0].max_staged}, max block {tiles[0].max_block_bytes} B, "
          f"{e / max(tiles[0].n_entries, 1):.2f} edges per shared-memory row read", flush=True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, generator=g).to(dev); gout = torch.randn(n, h, generator=g).to(dev)
    w, b, eps = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev), torch.zeros(1, device=dev)
    out = torch.empty_like(x)
    bf = 2 * n * h * 4 + e * 8 + (n + 1) * 4 + 8 * h
    bb = 3 * n * h * 4 + e * 8 + (n + 1) * 4
    for tiled in (False, True):
        tf = timeit(lambda: K.gine_aggr_fwd(x, sg, w, b, eps, out, tiled=tiled))
        tb = timeit(lambda: K.gine_aggr_bwd(gout, x, sg, w, b, eps, None, out, tiled=tiled))
        print(f"  {'tiled  ' if tiled else 'untiled'} fwd {tf:7.1f} us = {bf/tf/1e3:6.0f} GB/s ({bf/tf/1e3/6550.7*100:4.1f} %)   "
              f"bwd {tb:7.1f} us = {bb/tb/1e3:6.0f} GB/s ({bb/tb/1e3/6550.7*100:4.1f} %)", flush=True)


if which in ("c4", "all"):
    n = 100_000
    ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
    run("config 4", ei, ea, n)
if which == "c4x10":
    n = 1_000_000
    ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
    run("config 4 x 10", ei, ea, n)
if which in ("ref", "all"):
    coords = syn.station_coords(122, 600.0, 0)
    ei1, ea1 = G.radius_graph(syn.distance_matrix(coords), 100.0)
    for batch in (512, 4096):
        ei, ea = G.collate_static(ei1, ea1, 122, batch)
        run(f"reference graph x {batch}", ei, ea, 122 * batch)
