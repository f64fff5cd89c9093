"""Experiment: where does the reference-shape step time go?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from raincast_gnn_b200 import _lib, kernels as K
from raincast_gnn_b200.engine import TrainEngine
from raincast_gnn_b200.graph import build_station_graph
from raincast_gnn_b200.models import GNN

dev = torch.device("cuda:0")
bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ei, ea, ei_b, ea_b = B.static_graph(bsz)
m = bsz * B.N_STATIONS
sg = build_station_graph(ei_b, ea_b, m).to(dev)
model = B.seeded_model(GNN).to(dev).train()
eng = TrainEngine(model, sg, m, B.MEMBERS, B.FEATS, mode="graph").capture()
engp = TrainEngine(B.seeded_model(GNN).to(dev).train(), sg, m, B.MEMBERS, B.FEATS, mode="program").capture()
from raincast_gnn_b200.utils import synthetic as syn
x, ens = syn.node_features(m, B.MEMBERS, B.FEATS, seed=1); y = syn.log_precip_targets(m, seed=1)
eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
engp.load_batch(x.to(dev), ens.to(dev), y.to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=100, do_flush=False):
    for _ in range(5): fn()
    tot = 0.0
    if do_flush:
        for _ in range(n):
            flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); c.record(); c.synchronize(); tot += a.elapsed_time(c)
        return tot / n * 1e3
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); c.synchronize()
    return a.elapsed_time(c) / n * 1e3

print(f"B={bsz} launches/step {eng.launches_per_step}")
print(f"step, L2 flushed      : {timeit(eng.step, do_flush=True):8.1f} us")
print(f"step, back to back    : {timeit(eng.step):8.1f} us")
print(f"graph replay only     : {timeit(eng._graph.replay):8.1f} us")
print(f"optimizer only        : {timeit(eng._optimizer):8.1f} us")
print(f"program: ops {engp.prog_ops} phases {engp.prog_phases}")
print(f"program step, flushed : {timeit(engp.step, do_flush=True):8.1f} us")
print(f"program step, b2b     : {timeit(engp.step):8.1f} us")
# eager (no graph) for comparison
eng2 = TrainEngine(B.seeded_model(GNN).to(dev).train(), sg, m, B.MEMBERS, B.FEATS, use_cuda_graph=False, mode="graph").capture()
eng2.load_batch(x.to(dev), ens.to(dev), y.to(dev))
print(f"eager step (no graph) : {timeit(eng2.step, n=30):8.1f} us")
# forward only pieces, eager timing of individual blocks back to back (warm L2)
blk = eng._blocks
def fwd_ds(): K.deepsets_fwd(blk['ds'][0], eng.ens)
print(f"deepsets fwd eager    : {timeit(fwd_ds, n=50):8.1f} us")
emb, _ = K.deepsets_fwd(blk['ds'][0], eng.ens)
node, _ = K.dimred_fwd(blk['dr'][0], eng.x, emb)
def fwd_l(): K.gine_layer_fwd(blk['layers'][1][0], node, eng.graph, first=False, training=True)
print(f"gine layer fwd eager  : {timeit(fwd_l, n=50):8.1f} us")
# chain of trivial kernels in a graph: per-node overhead
z = torch.zeros(1, device=dev)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): z.add_(1)
    torch.cuda.current_stream().synchronize()
    with torch.cuda.graph(g):
        for _ in range(74): z.add_(1)
print(f"74 chained tiny nodes : {timeit(g.replay):8.1f} us")

# grid barrier cost: a program of 50 empty phases
import ctypes as C
L = _lib.lib()
L.rc_prog_begin()
for _ in range(50): L.rc_prog_nop()
nb = L.rc_prog_bytes(); buf = torch.empty(nb, dtype=torch.uint8, device=dev); info = (C.c_int * 4)()
_lib.check(L.rc_prog_end(buf.data_ptr(), nb, info))
st = torch.cuda.current_stream().cuda_stream
print(f"50 empty phases       : {timeit(lambda: L.rc_prog_run(buf.data_ptr(), info, st)):8.1f} us")
