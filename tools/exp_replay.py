"""Graph-replay time of the reference-shape step (B=8), back to back: python tools/exp_replay.py [label]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from raincast_gnn_b200.engine import TrainEngine
from raincast_gnn_b200.graph import build_station_graph
from raincast_gnn_b200.models import GNN
from raincast_gnn_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
ei, ea, ei_b, ea_b = B.static_graph(8)
m = 8 * B.N_STATIONS
sg = build_station_graph(ei_b, ea_b, m).to(dev)
eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), sg, m, B.MEMBERS, B.FEATS).capture()
x, ens = syn.node_features(m, B.MEMBERS, B.FEATS, seed=1); y = syn.log_precip_targets(m, seed=1)
eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
for _ in range(10): eng._graph.replay()
torch.cuda.synchronize()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200): eng._graph.replay()
c.record(); c.synchronize()
print(sys.argv[1] if len(sys.argv) > 1 else "", "graph replay %.1f us, launches %d" % (a.elapsed_time(c) / 200 * 1e3, eng.kernels_per_step))
