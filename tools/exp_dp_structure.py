"""Single GPU: does the data-parallel step's extra graph structure (a side-stream root kernel + an event the backward waits
for) cost time by itself?  Replays the captured step with and without it."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from raincast_gnn_b200 import kernels as K
from raincast_gnn_b200.engine import TrainEngine
from raincast_gnn_b200.graph import build_station_graph
from raincast_gnn_b200.models import GNN
from raincast_gnn_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
ei, ea, ei_b, ea_b = B.static_graph(8)
m = 8 * B.N_STATIONS
sg = build_station_graph(ei_b, ea_b, m).to(dev)
x, ens = syn.node_features(m, B.MEMBERS, B.FEATS, seed=1); y = syn.log_precip_targets(m, seed=1)
scratch = torch.zeros(32, device=dev)
for fake in (0, 1, 0, 1):
    eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), sg, m, B.MEMBERS, B.FEATS)
    if fake:
        ev = {}
        orig_ds, orig_hb = K.deepsets_fwd, K.head_bwd
        def ds(*a, **k):
            with K.on_side():
                scratch.add_(1.0)
                ev["e"] = torch.cuda.Event(); ev["e"].record(torch.cuda.current_stream(dev))
            return orig_ds(*a, **k)
        def hb(*a, **k):
            torch.cuda.current_stream(dev).wait_event(ev["e"])
            return orig_hb(*a, **k)
        K.deepsets_fwd, K.head_bwd = ds, hb
    eng.capture()
    if fake:
        K.deepsets_fwd, K.head_bwd = orig_ds, orig_hb
    eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
    for _ in range(10): eng._graph.replay()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(300): eng._graph.replay()
    c.record(); c.synchronize()
    print(f"side root kernel + event before backward: {fake}   {a.elapsed_time(c) / 300 * 1e3:.1f} us per replay")
