"""Per-step device time distribution of the captured reference-shape step (L2 flushed between steps, as bench.py does)."""
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
eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), build_station_graph(ei_b, ea_b, m).to(dev), m, B.MEMBERS, B.FEATS).capture()
x, ens = syn.node_features(m, B.MEMBERS, B.FEATS, seed=1); y = syn.log_precip_targets(m, seed=1)
eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in ("flush", "warm"):
    ev = []
    for i in range(420):
        if mode == "flush":
            flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng._graph.replay(); c.record(); ev.append((a, c))
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(c) * 1e3 for a, c in ev[20:]])
    q = torch.quantile(t, torch.tensor([0.05, 0.5, 0.9, 0.99]))
    # expected max of two / eight independent draws
    def emax(k):
        idx = torch.randint(0, t.numel(), (20000, k)); return t[idx].max(1).values.mean().item()
    print(f"{mode}: mean {t.mean():.1f} us  std {t.std():.1f}  p5 {q[0]:.1f} p50 {q[1]:.1f} p90 {q[2]:.1f} p99 {q[3]:.1f}  max {t.max():.1f};  E[max of 2] {emax(2):.1f}  E[max of 8] {emax(8):.1f}")
