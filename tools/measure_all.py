"""Secondary measurements of SURVEY.md 8(d) on one B200 (the headline line comes from bench.py):
batch sweep at the reference shape, fwd+bwd vs full step, config 4, config 5, CRPS bandwidth, and the CPU oracle
run eagerly ON the GPU (PyTorch library kernels) as the "existing Blackwell path" comparator.

    python tools/measure_all.py > profiles/r01_measurements.txt
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from raincast_gnn_b200 import graph as G, kernels as K  # noqa: E402
from raincast_gnn_b200.engine import TrainEngine  # noqa: E402
from raincast_gnn_b200.models import GNN  # noqa: E402
from raincast_gnn_b200.utils import synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    peaks = {"hbm_gbs": 6550.7}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n, do_flush=True, warm=3):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(n):
        if do_flush:
            flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        c.synchronize()
        tot += a.elapsed_time(c)
    return tot / n


print(f"device: {torch.cuda.get_device_name(0)}; HBM peak (measured copy) {peaks['hbm_gbs']} GB/s; "
      "all times CUDA events, L2 flushed between iterations\n")

# ---------------------------------------------------------------- reference shape: batch sweep, fwd+bwd vs full step
print("== reference shape (122 stations x 11 members x 35 features, H=128, L=4, mixed_u, fp32), one GPU ==")
print(f"{'B graphs':>9s} {'nodes':>8s} {'fwd+bwd ms':>11s} {'step ms':>9s} {'graphs/s':>10s} {'launches':>9s}")
for bsz in (8, 64, 512, 4096):
    ei, ea, ei_b, ea_b = B.static_graph(bsz)
    m = bsz * B.N_STATIONS
    sg = G.build_station_graph(ei_b, ea_b, m).to(dev)
    eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), sg, m, B.MEMBERS, B.FEATS).capture()
    x, ens = syn.node_features(m, B.MEMBERS, B.FEATS, seed=1)
    y = syn.log_precip_targets(m, seed=1)
    eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
    n = 50 if bsz <= 64 else 10
    t_fb = timed(eng._graph.replay, n)
    t_step = timed(eng.step, n)
    print(f"{bsz:9d} {m:8d} {t_fb:11.3f} {t_step:9.3f} {bsz / t_step * 1e3:10.0f} {eng.launches_per_step:9d}")
    del eng
    torch.cuda.empty_cache()

# ---------------------------------------------------------------- the CPU oracle run eagerly on the GPU (library kernels)
print("\n== comparator: the oracle's plain PyTorch ops run eagerly on the same B200 (cuBLAS / ATen kernels), B=8 ==")
from oracle import model as om, pyg as opyg  # noqa: E402
ei, ea, ei_b, ea_b = B.static_graph(8)
ref = B.seeded_model(om.GNN).to(dev).train()
opt = torch.optim.AdamW(ref.parameters(), lr=1e-4)
x, ens = syn.node_features(8 * B.N_STATIONS, B.MEMBERS, B.FEATS, seed=1)
y = syn.log_precip_targets(8 * B.N_STATIONS, seed=1)
d = opyg.Data(x=x.to(dev), ensemble=ens.to(dev), edge_index=ei_b.to(dev), edge_attr=ea_b.to(dev), y=y.to(dev))


def eager():
    loss = ref.loss_fn.crps(ref(d), d.y)
    opt.zero_grad()
    loss.backward()
    opt.step()


t = timed(eager, 20, do_flush=False)
print(f"eager PyTorch step on B200: {t:.3f} ms -> {8 / t * 1e3:.0f} graphs/s")


# ---------------------------------------------------------------- config 4 / config 5
def big_step(hidden, em, bf16, n=100_000):
    ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
    sg = G.build_station_graph(ei, ea, n).to(dev)
    kw = dict(B.MODEL_KW, hidden_channels_gnn=hidden, out_channels_gnn=hidden)
    model = GNN(**kw)
    model.load_state_dict(syn.seeded_state_dict(model.state_dict(), seed=2024))
    model.to(dev).train()
    if bf16:
        model.deepset.compute_dtype = "bf16"
    eng = TrainEngine(model, sg, n, em, B.FEATS, use_cuda_graph=False)
    x, ens = syn.node_features(n, em, B.FEATS, seed=3)
    y = syn.log_precip_targets(n, seed=3)
    eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
    t = timed(eng.step, 5, warm=2)
    mem = torch.cuda.max_memory_allocated() / 2**30
    del eng, model
    torch.cuda.empty_cache()
    return t, mem


print("\n== scaled graph: 100 000 nodes, 2 978 560 edges, 51 members, one graph per step ==")
t, mem = big_step(128, 51, False)
print(f"config 4 (H=128, fp32; DeepSets forward on tcgen05 3xTF32): {t:.2f} ms/step, peak {mem:.2f} GiB   "
      "(CPU oracle, 16 threads: ~4.2 s fwd+bwd)")
t, mem = big_step(512, 51, True)
print(f"config 5 (H=512, bf16 DeepSets member contraction on tcgen05, GINE fp32): {t:.2f} ms/step, peak {mem:.2f} GiB")

# ---------------------------------------------------------------- CRPS bandwidth
print("\n== links + mixture CRPS value + gradient, one pass (algorithmic bytes (2C+1)*4 = 44 B/node at C=5) ==")
for mm in (1 << 20, 1 << 24):
    raw = torch.randn(mm, 5, device=dev)
    y = syn.log_precip_targets(mm, seed=5).to(dev)
    t = timed(lambda: K.crps_fwd_bwd(raw, y, 3, raw_input=True), 10)
    gbs = 44 * mm / t / 1e6
    print(f"M = 2^{mm.bit_length() - 1}: {t * 1e3:8.1f} us  -> {gbs:7.0f} GB/s = {gbs / peaks['hbm_gbs'] * 100:4.1f} % of measured HBM peak")

# ---------------------------------------------------------------- aggregation
res, bf, bb, n, e = B.measure_aggregation(dev)
print(f"\n== GINE aggregation, config-4 graph, H=128 ==\nfwd {res['fwd'] * 1e3:.1f} us = {bf / res['fwd'] / 1e6:.0f} GB/s "
      f"({bf / res['fwd'] / 1e6 / peaks['hbm_gbs'] * 100:.1f} %)   bwd {res['bwd'] * 1e3:.1f} us = {bb / res['bwd'] / 1e6:.0f} GB/s "
      f"({bb / res['bwd'] / 1e6 / peaks['hbm_gbs'] * 100:.1f} %)")
