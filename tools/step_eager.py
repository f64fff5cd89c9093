"""A few eager (un-captured) reference-shape steps, for `ncu --cache-control none` launch lists: python tools/step_eager.py [steps]"""
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
eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), sg, m, B.MEMBERS, B.FEATS, use_cuda_graph=False)
x, ens = syn.node_features(m, B.MEMBERS, B.FEATS, seed=1); y = syn.log_precip_targets(m, seed=1)
eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    eng.step()
torch.cuda.synchronize()
print("done")
