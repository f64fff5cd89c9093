#!/usr/bin/env python3
"""bench.py — station-graphs/s of the fused GINE + DeepSets + mixture-CRPS training step on B200.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                   (CPU arm: the oracle port on the host cores)

One "step" = one train.py iteration (train.py:61-71) on one batch of B=8 station graphs per GPU of the
reference shape (122 stations x 11 members x 35 features, H=128, 4 GINE layers, MixedLoss with learned u,
fp32): forward, CRPS, backward, gradient all-reduce (N > 1), AdamW.  Prints ONE JSON line on rank 0.

  value  — whole-job graphs/s with the batch already resident in HBM (CUDA-event time of the K steps)
  e2e    — the same through the public engine calls with HOST (pinned) batches: inside every timed step one H2D
           copy of x / ensemble / y (the next step's batch, prefetched on a copy stream while this step runs, as
           train.py does) and a D2H read of the loss
  roofline — the GINE aggregation forward kernel on the config-4 graph (100k nodes, 2 978 560 edges,
           H=128): algorithmic bytes 2*M*H*4 + E*8 + (M+1)*4 + 8*H   / mean CUDA-event time;
           `bwd` and `batched_reference_graphs` (4096 reference graphs in one batch) report the same for the
           backward kernel and for the graph shape where HBM, not the SM, is the relevant bound
           `deepsets_contraction`: the tcgen05 member contraction at the config-4 / config-5 shapes (TFLOP/s issued
           against the measured dense peak, ensemble read rate against the HBM peak)
           `crps`: the links + mixture-CRPS kernel at 2^24 nodes against (2C+1)*4 bytes per node
  step_roofline — FLOPs and bytes of one reference-shape step against the measured fp32 FMA peak (rc_debug_fma_peak)
           and the HBM peak: how far the launch-bound step is from either floor
  gpu_eager_baseline — the oracle's modules (plain PyTorch ops, library kernels) run eagerly on the same GPU
  config4 / config5 — whole training step (ms) of one 100k-node graph, 51 members: H=128 fp32 / H=512 bf16 DeepSets
  train_loop_e2e — raincast_gnn_b200.train.run_epoch_engine over a shuffled DataLoader (host collate included), wall clock
  train_loop_resident — raincast_gnn_b200.train.run_epoch_resident (train.py's default loop: the split lives in HBM), wall clock
  dp_check — data-parallel semantics: the engine's emulated-rank step against the micro-batch oracle (N = 1), replicas
           bit-identical and losses equal to the NCCL exchange (N > 1)
  cpu_baseline — the CPU oracle (reference modules' arithmetic, oracle/) timed on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# NCCL prints its version banner on stdout; the contract is ONE JSON line there
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

import torch  # noqa: E402

B_PER_GPU, N_STATIONS, MEMBERS, FEATS, HIDDEN, LAYERS = 8, 122, 11, 35, 128, 4
WORKLOAD = ("24h_mixed_u reference shape: B=8 graphs/GPU x 122 stations x 11 members x F=35, H=128, L=4, "
            "MixedLoss(grad_u=True, xi=0.5), AdamW lr 1e-4, fp32; synthetic data (SURVEY.md 8d)")
MODEL_KW = dict(in_channels=FEATS, hidden_channels_gnn=HIDDEN, out_channels_gnn=HIDDEN, num_layers_gnn=LAYERS,
                optimizer_class=torch.optim.AdamW, optimizer_params={"lr": 1e-4}, loss="MixedLoss", grad_u="True",
                u=1.71, xi=0.5)





def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-roofline", action="store_true", help="skip the config-4 aggregation measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-sweep", action="store_true", help="also report graphs/s at B in {64, 512, 4096} per GPU (SURVEY.md 8d)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- data / model
def make_host_batches(n_batches: int, batch: int, seed: int, rank: int, world: int):
    """Pinned host batches of `batch` dates each; dates are sharded rank::world of a seeded permutation."""
    from raincast_gnn_b200 import dp
    from raincast_gnn_b200.utils import synthetic as syn
    n_dates = n_batches * batch * world
    mine = dp.shard_dates(n_dates, rank, world, seed=seed)
    m = batch * N_STATIONS
    out = []
    for i in range(n_batches):
        dates = mine[i * batch:(i + 1) * batch]
        s = 1000 + dates[0]                               # a date-dependent seed: different ranks see different data
        x, ens = syn.node_features(m, MEMBERS, FEATS, seed=s)
        y = syn.log_precip_targets(m, seed=s)
        out.append(tuple(t.pin_memory() if torch.cuda.is_available() else t for t in (x, ens, y)))
    return out


def static_graph(batch: int, oracle_only: bool = False):
    """The synthetic reference station graph and its `batch`-fold collation.  oracle_only: built with
    oracle/graph.py (numpy) so that the CPU arm never loads the repo's native library."""
    from raincast_gnn_b200.utils import synthetic as syn
    dist = syn.distance_matrix(syn.station_coords(N_STATIONS, 600.0, seed=0))
    if oracle_only:
        import numpy as np
        from oracle import graph as og
        ei, ea = og.radius_graph(np.asarray(dist), 100.0)
        ei_b, ea_b = og.collate_edges(ei, ea, N_STATIONS, batch)
        ei, ea, ei_b, ea_b = (torch.as_tensor(t) for t in (ei, ea, ei_b, ea_b))
        ea, ea_b = ea.reshape(-1, 1), ea_b.reshape(-1, 1)
    else:
        from raincast_gnn_b200 import graph as G
        ei, ea = G.radius_graph(dist, 100.0)
        ei_b, ea_b = G.collate_static(ei, ea, N_STATIONS, batch)
    assert ei.shape[1] == 1164, "synthetic reference graph must have 1 164 edges (SURVEY.md 8d)"
    return ei, ea, ei_b, ea_b


def seeded_model(ctor):
    from raincast_gnn_b200.utils import synthetic as syn
    model = ctor(**MODEL_KW)
    model.load_state_dict(syn.seeded_state_dict(model.state_dict(), seed=2024))
    return model


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_steps_per_second(steps: int, warmup: int, budget_s: float):
    """Oracle port (reference arithmetic in plain CPU torch) training steps on all host cores."""
    from oracle import model as om, pyg as opyg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = seeded_model(om.GNN)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    ei, ea, ei_b, ea_b = static_graph(B_PER_GPU, oracle_only=True)
    batches = make_host_batches(4, B_PER_GPU, seed=7, rank=0, world=1)

    def one(i):
        x, ens, y = batches[i % len(batches)]
        d = opyg.Data(x=x, ensemble=ens, edge_index=ei_b, edge_attr=ea_b, y=y)
        loss = model.loss_fn.crps(model(d), d.y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.item()
    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    while done < steps and (time.perf_counter() - t0) < budget_s:
        one(done)
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, done, dt, cores


def run_reference(args):
    rank, _, world = __import__("raincast_gnn_b200.dp", fromlist=["env_world"]).env_world()
    if rank != 0:
        return
    sps, done, dt, cores = cpu_steps_per_second(args.steps, max(args.warmup, 3), budget_s=150.0)
    value = sps * B_PER_GPU
    line = {"impl": "reference", "metric": "station-graphs/sec train (fwd+bwd+CRPS+AdamW)", "value": value, "unit": "graphs/s",
            "n_gpus": args.gpus, "steps": done, "warmup": max(args.warmup, 3), "ms_per_step": 1000.0 * dt / max(done, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU},
            "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": cores, "kind": "port",
                             "sample": f"{done} train steps of B=8 graphs (oracle port of models/*.py + PyG GINEConv restatement), "
                                       f"torch.set_num_threads({cores})"},
            "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- roofline leg
def measure_aggregation(dev, iters: int = 20, which: str = "config4"):
    """GINE aggregation fwd (and bwd), L2 flushed between launches so x comes from HBM.
    which = "config4": the 100k-node radius graph of SURVEY.md 8d (30 edges per row);
    which = "ref4096": 4096 reference station graphs in one batch (499 712 nodes, 9.5 edges per row, block diagonal)."""
    from raincast_gnn_b200 import graph as G
    from raincast_gnn_b200 import kernels as K
    from raincast_gnn_b200.utils import synthetic as syn
    h = HIDDEN
    if which == "config4":
        n = 100_000
        ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
    else:
        ei1, ea1, _, _ = static_graph(1)
        n = 4096 * N_STATIONS
        ei, ea = G.collate_static(ei1, ea1, N_STATIONS, 4096)
    sg = G.build_station_graph(ei, ea, n).to(dev)
    e = ei.shape[1]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, generator=g).to(dev)
    gout = torch.randn(n, h, generator=g).to(dev)
    w, b, eps = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev), torch.zeros(1, device=dev)
    out = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sg.tiles(h)                    # the product path: station tiles (built once per graph, like the CSR)

    def fwd():
        K.gine_aggr_fwd(x, sg, w, b, eps, out)

    def bwd():
        K.gine_aggr_bwd(gout, x, sg, w, b, eps, None, out)
    res = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        times = []
        for _ in range(iters):
            flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            c.record()
            c.synchronize()
            times.append(a.elapsed_time(c))
        res[name] = sum(times) / len(times)
    bytes_fwd = 2 * n * h * 4 + e * 8 + (n + 1) * 4 + 8 * h
    bytes_bwd = 3 * n * h * 4 + e * 8 + (n + 1) * 4
    return res, bytes_fwd, bytes_bwd, n, e


def measure_deepsets_contraction(dev, peaks, iters: int = 5):
    """The tcgen05 member contraction (Linear + ReLU + sum over members) at the config-4 shape (fp32-accurate 3xTF32) and
    the config-5 shape (bf16 operands, H = 512): CUDA-event time per launch, algorithmic TFLOP/s (2*M*Em*F*H), the tensor
    work actually issued against the measured dense peak (TF32 = half the bf16 rate, three products per fp32 product,
    K padded to the MMA's k-block), and the ensemble read rate against the HBM peak."""
    from raincast_gnn_b200 import _lib
    L = _lib.lib()
    st = torch.cuda.current_stream(dev).cuda_stream
    m, em, f = 100_000, 51, FEATS
    g = torch.Generator().manual_seed(0)
    ens = torch.randn(m, em, f, generator=g).to(dev)
    out = {}
    for tag, h, bf16 in (("config4_fp32_3xtf32", HIDDEN, False), ("config5_bf16_h512", 512, True)):
        w1 = (torch.randn(h, f, generator=g) * 0.2).to(dev)
        b1 = torch.randn(h, generator=g).to(dev)
        pooled = torch.empty(m, h, device=dev)
        fn = L.rc_deepsets_pool_fwd_bf16 if bf16 else L.rc_deepsets_pool_fwd
        run = lambda: _lib.check(fn(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), pooled.data_ptr(), m, em, f, h, st))
        for _ in range(2):
            run()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            run()
        c.record()
        c.synchronize()
        ms = a.elapsed_time(c) / iters                      # the 714 MB ensemble is larger than L2: every launch reads HBM
        kpad = (f + 15) // 16 * 16 if bf16 else (f + 7) // 8 * 8
        issued = 2.0 * (m * em) * kpad * h * (1 if bf16 else 3)
        peak_tf = peaks.get("bf16_tflops", 1686.5) * (1.0 if bf16 else 0.5)
        out[tag] = {"workload": f"{m} stations x {em} members x {f} features -> H={h}", "us_per_launch": ms * 1e3,
                    "tflops_algorithmic": 2.0 * m * em * f * h / ms / 1e9, "tflops_issued": issued / ms / 1e9,
                    "tensor_peak_tflops": peak_tf, "frac_tensor": issued / ms / 1e9 / peak_tf,
                    "ens_read_gbs": ens.numel() * 4 / ms / 1e6, "frac_hbm": ens.numel() * 4 / ms / 1e6 / peaks["hbm_gbs"]}
        del w1, b1, pooled
    return out


def measure_crps(dev, peaks, log2_m: int = 24, iters: int = 5):
    """The links + mixture-CRPS kernel (value + gradient, MixedLoss with learned u, raw head input) at M = 2^24 nodes:
    whole rc_crps_fwd_bwd call (count + main + final kernels) against (2C+1)*4 algorithmic bytes per node."""
    from raincast_gnn_b200 import kernels as K
    from raincast_gnn_b200.utils import synthetic as syn
    mm, width = 1 << log2_m, 5
    raw = torch.randn(mm, width, device=dev)
    y = syn.log_precip_targets(mm, seed=5).to(dev)
    for _ in range(3):
        K.crps_fwd_bwd(raw, y, 3, raw_input=True)
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        K.crps_fwd_bwd(raw, y, 3, raw_input=True)
    c.record()
    c.synchronize()
    ms = a.elapsed_time(c) / iters
    by = (2 * width + 1) * 4 * mm
    return {"workload": f"MixedLoss(grad_u) value + gradient, raw head input, M = 2^{log2_m} nodes (inputs 11x L2: no flush needed)",
            "us_per_launch": ms * 1e3, "algorithmic_bytes": by, "achieved": by / ms / 1e6, "unit": "GB/s",
            "frac": by / ms / 1e6 / peaks["hbm_gbs"]}


def measure_fp32_peak(dev):
    """fp32 FMA throughput of this GPU (MEASURED_PEAKS.json has no fp32 figure): 16 independent FFMA chains per thread,
    8 CTAs of 256 threads per SM; the best of five launches of ~1 ms."""
    import ctypes as C
    from raincast_gnn_b200 import _lib
    L = _lib.lib()
    scratch = torch.zeros(4, device=dev)
    flops = C.c_double(0.0)
    st = torch.cuda.current_stream(dev).cuda_stream
    best = 0.0
    for _ in range(6):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.rc_debug_fma_peak(scratch.data_ptr(), 2000, C.byref(flops), st), "rc_debug_fma_peak")
        c.record()
        c.synchronize()
        best = max(best, flops.value / (a.elapsed_time(c) * 1e-3) / 1e12)
    return best


def step_flops_bytes():
    """Algorithmic work of one reference-shape training step (hoisted DeepSets form: the second phi Linear is
    applied once per station): forward FLOPs x 3 (backward = data + weight gradients), and the compulsory HBM bytes - the
    batch in, every parameter / gradient / Adam moment once (28 B per parameter), every saved activation written and read once."""
    m, em, f, h, c, L = B_PER_GPU * N_STATIONS, MEMBERS, FEATS, HIDDEN, 5, LAYERS
    e = 1164 * B_PER_GPU
    fwd = 2 * m * em * f * h + 3 * 2 * m * h * h + 2 * m * (f + h) * h + L * (2 * 2 * m * h * h + 3 * e * h) + 2 * m * h * c
    params = h * f + h + 3 * (h * h + h) + h * (f + h) + h + L * (1 + 2 * h + 2 * (h * h + h) + 2 * h) + c * h + c
    acts = m * h * (4 + 1 + L * 4)                 # pooled, s2, r1, emb, node; per layer x, agg, t, y
    bytes_ = 4 * (m * f + m * em * f + m) + 28 * params + 2 * 4 * acts + 2 * 8 * e
    return 3.0 * fwd, float(bytes_), params


def measure_gpu_eager(dev, steps=20, warmup=5):
    """The oracle's plain PyTorch ops run eagerly on this GPU (library kernels: the only 'existing Blackwell path' the
    reference has): train steps per second on the benchmark's batch, CUDA events."""
    from oracle import model as om, pyg as opyg
    model = seeded_model(om.GNN).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    ei, ea, ei_b, ea_b = static_graph(B_PER_GPU, oracle_only=True)
    batches = make_host_batches(2, B_PER_GPU, seed=7, rank=0, world=1)
    dbs = [opyg.Data(x=x.to(dev), ensemble=en.to(dev), edge_index=ei_b.to(dev), edge_attr=ea_b.to(dev), y=y.to(dev)) for x, en, y in batches]

    def one(i):
        d = dbs[i % len(dbs)]
        loss = model.loss_fn.crps(model(d), d.y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss
    for i in range(warmup):
        one(i)
    torch.cuda.synchronize(dev)
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        one(i).item()                                   # train.py:71 reads the loss every step
    c.record()
    c.synchronize()
    ms = a.elapsed_time(c) / steps
    return {"value": B_PER_GPU / (ms * 1e-3), "unit": "graphs/s", "ms_per_step": ms,
            "what": "oracle modules (reference arithmetic, PyG GINEConv restatement) in eager PyTorch on this GPU, library kernels, fp32"}


def measure_scaled_step(dev, hidden: int, bf16: bool, iters: int = 3):
    """One whole training step at the config-4 shape (one 100k-node graph, 2 978 560 edges, 51 members): ms per step,
    CUDA events, kernels issued eagerly (a step of several ms is not launch bound)."""
    from raincast_gnn_b200 import graph as G
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.utils import synthetic as syn
    n, em = 100_000, 51
    ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
    sg = G.build_station_graph(ei, ea, n).to(dev)
    x, ens = syn.node_features(n, em, FEATS, seed=3)
    y = syn.log_precip_targets(n, seed=3)
    kw = dict(MODEL_KW, hidden_channels_gnn=hidden, out_channels_gnn=hidden)
    model = GNN(**kw)
    model.load_state_dict(syn.seeded_state_dict(model.state_dict(), seed=2024))
    model.to(dev).train()
    if bf16:
        model.deepset.compute_dtype = "bf16"
    eng = TrainEngine(model, sg, n, em, FEATS, use_cuda_graph=False)
    eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
    for _ in range(2):
        eng.step()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        eng.step()
    c.record()
    c.synchronize()
    out = {"ms_per_step": a.elapsed_time(c) / iters, "loss": float(eng.loss.item())}
    del eng, model, sg
    torch.cuda.empty_cache()
    return out


def measure_refshape_step(dev, b: int, hidden: int = HIDDEN, bf16: bool = False, flush=None):
    """The training step on `b` collated reference station graphs: ms per step (CUDA events per step, L2 flushed between
    steps when `flush` is given).  One CUDA graph while the step is launch bound (M < 16k rows), eager beyond."""
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.graph import build_station_graph
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.utils import synthetic as syn
    _, _, ei_b, ea_b = static_graph(b)
    m = b * N_STATIONS
    sg = build_station_graph(ei_b, ea_b, m).to(dev)
    kw = dict(MODEL_KW, hidden_channels_gnn=hidden, out_channels_gnn=hidden)
    model = GNN(**kw)
    model.load_state_dict(syn.seeded_state_dict(model.state_dict(), seed=2024))
    model.to(dev).train()
    if bf16:
        model.deepset.compute_dtype = "bf16"
    graph = m < 16384
    eng = TrainEngine(model, sg, m, MEMBERS, FEATS, lr=1e-4, use_cuda_graph=graph)
    if graph:
        eng.capture()
    x, ens = syn.node_features(m, MEMBERS, FEATS, seed=11)
    y = syn.log_precip_targets(m, seed=11)
    eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
    iters = 20 if b <= 512 else 5
    for _ in range(3):
        eng.step()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.step()
        c.record()
        c.synchronize()
        total += a.elapsed_time(c)
    ms = total / iters
    out = {"ms_per_step": ms, "graphs_per_s": b / (ms * 1e-3), "nodes": m, "edges": int(ei_b.shape[1]),
           "cuda_graph": graph, "launches_per_step": eng.launches_per_step, "loss": float(eng.loss.item())}
    del eng, model, sg, x, ens, y
    torch.cuda.empty_cache()
    return out


def measure_batch_sweep(dev, batches=(64, 512, 4096)):
    """The reference-shape training step at larger per-GPU batches (SURVEY.md 8d's sweep): the same model and station
    graph, B dates collated into one batch.  B = 8 is launch / dependency latency; from a few hundred graphs on the step
    is throughput bound (M >= 16k rows: tensor-core Linears, tiled aggregation).  L2 flushed between steps up to
    B = 512 (the B = 4096 inputs alone are 7x L2)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    return {str(b): measure_refshape_step(dev, b, flush=flush if b <= 512 else None) for b in batches}


def measure_train_loop(dev, n_dates: int = 512):
    """train.py's epoch loop as shipped (run_epoch_engine): DataLoader collate on the host, pinning, the next batch's
    H2D copy prefetched on a copy stream, one captured step per batch, the ragged last batch stepped eagerly, one loss
    read per epoch.  graphs/s over one epoch of `n_dates` synthetic dates (wall clock, after a warm-up epoch)."""
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import DataLoader
    from raincast_gnn_b200.train import run_epoch_engine
    from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
    ds = SyntheticEUPPBench(n_dates=n_dates + 3, members=MEMBERS)       # + 3: the epoch ends on a ragged batch, like the reference's splits
    loader = DataLoader(ds, batch_size=B_PER_GPU, shuffle=True)
    first = next(iter(loader))
    model = seeded_model(GNN).to(dev).train()
    eng = TrainEngine(model, first.station_graph, first.x.shape[0], MEMBERS, FEATS, lr=1e-4).capture()
    run_epoch_engine(eng, loader)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    loss = run_epoch_engine(eng, loader)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return {"value": len(ds) / dt, "unit": "graphs/s", "epoch_s": dt, "dates": len(ds), "mean_loss": loss,
            "what": "run_epoch_engine over a shuffled DataLoader: host collate + pin + prefetched H2D + captured step; wall clock"}


def measure_train_loop_resident(dev, n_dates: int = 1024):
    """train.py's default epoch loop (run_epoch_resident): the split lives in HBM, the host draws the epoch's order and
    uploads it once, every step is one graph replay (rc_gather_dates_step + the captured step), the ragged last batch is
    stepped eagerly, one loss read per epoch."""
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import DataLoader
    from raincast_gnn_b200.train import run_epoch_resident
    from raincast_gnn_b200.utils.dataset import DeviceSplit, SyntheticEUPPBench
    ds = SyntheticEUPPBench(n_dates=n_dates + 3, members=MEMBERS)
    first = next(iter(DataLoader(ds, batch_size=B_PER_GPU)))
    model = seeded_model(GNN).to(dev).train()
    eng = TrainEngine(model, first.station_graph, first.x.shape[0], MEMBERS, FEATS, lr=1e-4).capture()
    split = DeviceSplit([ds[i] for i in range(len(ds))], dev)
    run_epoch_resident(eng, split, B_PER_GPU)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    loss = run_epoch_resident(eng, split, B_PER_GPU)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return {"value": len(ds) / dt, "unit": "graphs/s", "epoch_s": dt, "dates": len(ds), "mean_loss": loss,
            "what": "run_epoch_resident (train.py's default): split and the epoch's order resident in HBM, ONE graph replay per step "
                    "(gather of the next batch + the captured step), ragged last batch stepped eagerly; wall clock"}


def dp_check(dev, pg, rank, world, steps: int = 4):
    """Data-parallel correctness carried by the bench line (N > 1): after `steps` steps from one initialisation the
    replicas are bit-identical (max |difference| of the flat parameters across ranks = 0), and the peer-memory exchange
    gives the loss trajectory of the NCCL all-reduce (max relative difference)."""
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.graph import build_station_graph
    from raincast_gnn_b200.models import GNN
    ei, ea, ei_b, ea_b = static_graph(B_PER_GPU)
    m = B_PER_GPU * N_STATIONS
    batches = make_host_batches(steps, B_PER_GPU, seed=11, rank=rank, world=world)
    out = {}
    traj = {}
    for mode in ("p2p", "nccl"):
        os.environ["RC_DP_EXCHANGE"] = mode
        model = seeded_model(GNN).to(dev).train()
        eng = TrainEngine(model, build_station_graph(ei_b, ea_b, m).to(dev), m, MEMBERS, FEATS, lr=1e-3, process_group=pg).capture()
        ls = []
        for i in range(steps):
            eng.load_batch(*batches[i])
            ls.append(eng.step().clone())
        torch.cuda.synchronize(dev)
        traj[mode] = torch.cat(ls)
        flat = eng.flat_p.detach().clone()
        lo, hi = flat.clone(), flat.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN, group=pg)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX, group=pg)
        out[f"replicas_max_abs_diff_{mode}"] = float((hi - lo).abs().max())
        out[f"exchange_{mode}"] = "peer memory" if eng.p2p is not None else "nccl all-reduce"
        del eng, model
    os.environ.pop("RC_DP_EXCHANGE", None)
    out["replicas_identical"] = out["replicas_max_abs_diff_p2p"] == 0.0 and out["replicas_max_abs_diff_nccl"] == 0.0
    out["loss_vs_nccl"] = float(((traj["p2p"] - traj["nccl"]).abs() / traj["nccl"].abs()).max())
    return out


def dp_check_one_gpu(dev, ranks: int = 4, per_rank: int = 2):
    """N = 1: data-parallel SEMANTICS carried by the bench line without a second GPU.  `ranks` micro-batches of `per_rank`
    small station graphs go through the engine's emulated-rank step (per-rank BatchNorm statistics and valid-node means,
    gradients summed in rank order, ONE AdamW step with grad_scale = 1/ranks: the arithmetic of rc_p2p_step) and through
    the float64 oracle (independent passes from the same weights, mean of the gradients); reported: the worst relative
    difference of the per-rank losses and of the mean gradient tensors (max-norm; the bias in front of BatchNorm, whose
    true gradient is zero, on its weight's scale)."""
    from oracle import model as om, pyg as opyg
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import DataLoader
    from raincast_gnn_b200.utils import synthetic as syn
    from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
    ds = SyntheticEUPPBench(n_dates=ranks * per_rank, num_stations=40, members=11, feats=9, max_dist=200.0)
    loader = list(DataLoader(ds, batch_size=per_rank))
    kw = dict(MODEL_KW, in_channels=9, num_layers_gnn=2, optimizer_params={"lr": 1e-3})
    ours = GNN(**kw)
    sd = syn.seeded_state_dict(ours.state_dict(), seed=7)
    ours.load_state_dict(sd)
    ours.to(dev).train()
    eng = TrainEngine(ours, loader[0].station_graph, loader[0].x.shape[0], 11, 9, lr=1e-3, use_cuda_graph=False)
    losses, mean_grads = eng.step_emulated_ranks([(b.x, b.ensemble, b.y) for b in loader])
    ref = om.GNN(**kw).double()
    ref.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()})
    ref.conv.force_float = False
    ref.train()
    acc = {k: torch.zeros_like(p) for k, p in ref.named_parameters()}
    ref_losses = []
    for b in loader:
        d = opyg.Data(x=b.x.double(), ensemble=b.ensemble.double(), edge_index=b.edge_index, edge_attr=b.edge_attr.double(), y=b.y.double())
        ref.zero_grad()
        loss = ref.loss_fn.crps(ref(d), d.y)
        loss.backward()
        ref_losses.append(float(loss.detach()))
        for k, p in ref.named_parameters():
            acc[k] += p.grad
    worst = 0.0
    for k in acc:
        want = acc[k] / ranks
        scale = (acc[k[:-4] + "weight"] / ranks).abs().max().item() if k.endswith(".nn.0.bias") else want.abs().max().item()
        worst = max(worst, (mean_grads[k].cpu().double() - want).abs().max().item() / max(scale, 1e-12))
    lo = torch.tensor(ref_losses, dtype=torch.float64)
    return {"mode": f"{ranks} ranks emulated on one GPU vs the float64 micro-batch oracle", "loss_max_rel_err": float(((losses.cpu() - lo).abs() / lo.abs()).max()),
            "mean_gradient_max_rel_err": worst, "within_1e-5": bool(worst < 1e-5)}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


# --------------------------------------------------------------------------------------------- main arm
def run_b200(args):
    from raincast_gnn_b200 import _lib, dp
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.graph import build_station_graph
    from raincast_gnn_b200.models import GNN
    rank, local_rank, world = dp.env_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = dp.init_from_env("nccl")
    K_steps, W = args.steps, max(args.warmup, 3)
    ei, ea, ei_b, ea_b = static_graph(B_PER_GPU)
    m = B_PER_GPU * N_STATIONS
    sg = build_station_graph(ei_b, ea_b, m).to(dev)
    model = seeded_model(GNN).to(dev).train()
    eng = TrainEngine(model, sg, m, MEMBERS, FEATS, lr=1e-4, process_group=pg).capture()
    batches = make_host_batches(16, B_PER_GPU, seed=7, rank=rank, world=world)
    losses_host = torch.zeros(K_steps + W, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    align = torch.zeros(1, device=dev)
    h2d = sum(t.numel() * t.element_size() for t in batches[0])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(e2e: bool):
        """K steps, each bracketed by its own CUDA events (the L2 flush between steps is outside the events).
        e2e: every timed interval contains one complete host -> device copy of a batch (the NEXT step's, started
        right after the start event on the copy stream and waited for before the end event - the engine's input
        prefetch, as train.py uses it), the move of the current batch into the step's inputs, the step, and the
        device -> host read of its loss."""
        total_ms = 0.0
        if e2e:
            eng.prefetch(*batches[0])
        for i in range(W + K_steps):
            flush.zero_()
            if world > 1:
                # the un-timed L2 flush lets the ranks drift apart; without re-aligning them here (device side, no host
                # sync) the drift would be measured as part of the next step's gradient exchange
                eng.align_ranks()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if e2e:
                eng.take_prefetched()
                eng.prefetch(*batches[(i + 1) % len(batches)])
            loss = eng.step()
            if e2e:
                losses_host[i].copy_(loss[0], non_blocking=True)
                eng.wait_prefetch()
            c.record()
            if i >= W:
                c.synchronize()
                total_ms += a.elapsed_time(c)
        return total_ms

    eng.load_batch(*batches[0])
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_wall = time.perf_counter()
        ms_dev = timed_loop(e2e=False)
        barrier()
        wall_dev = time.perf_counter() - t_wall
        ms_e2e = timed_loop(e2e=True)
        barrier()
    eng.check_peers()
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    graphs = K_steps * B_PER_GPU * world
    value = graphs / (ms_dev / 1000.0)
    e2e_value = graphs / (ms_e2e / 1000.0)
    final_loss = float(losses_host[-1])

    line = {"metric": "station-graphs/sec train (fwd+bwd+CRPS+AdamW)", "value": value, "unit": "graphs/s", "n_gpus": world,
            "steps": K_steps, "warmup": W, "ms_per_step": ms_dev / K_steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU * world, "parallelism": (f"dp{world} (dates sharded; per step one exchange of {eng.n_params} fp32 gradients: " +
                                       ("waited for, summed from NVLink peer memory and applied by ONE kernel inside the captured step)" if eng.p2p is not None else "one NCCL all-reduce)"))
                                      if world > 1 else "single GPU (dates would be sharded rank::world; no gradient exchange)",
                       "timing": "per-step CUDA events on the launch stream; 256 MiB L2 flush between timed steps, outside the events (N > 1: the ranks are re-aligned by a device-side barrier over peer-memory flags after the flush, also outside the events); max over ranks",
                       "cuda_graph": True, "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / K_steps},
            "gpu_launches": eng.launches_per_step * K_steps, "launches_per_step": eng.launches_per_step,
            "clocks": clocks.summary(), "wall_s_device_loop": wall_dev}

    if world > 1:
        try:
            line["dp_check"] = dp_check(dev, pg, rank, world)
        except Exception as exc:
            line["dp_check"] = {"error": repr(exc)}
    else:
        try:
            line["dp_check"] = dp_check_one_gpu(dev)
        except Exception as exc:
            line["dp_check"] = {"error": repr(exc)}
    if rank == 0:
        peaks, peak_kind = load_peaks()
        if not args.no_roofline:
            res, bf, bb, n, e = measure_aggregation(dev)
            ach = bf / (res["fwd"] * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "kernel": "gine_aggr_fwd_tiled_kernel (config 4: 100k nodes, 2 978 560 edges, H=128; station tiles, L2 flushed before every launch)",
                                "achieved": ach, "peak": peaks["hbm_gbs"], "peak_kind": f"{peak_kind} hbm_gbs (burst copy)",
                                "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                                "algorithmic_bytes": bf, "us_per_launch": res["fwd"] * 1e3,
                                "bwd": {"achieved": bb / (res["bwd"] * 1e-3) / 1e9, "algorithmic_bytes": bb,
                                        "us_per_launch": res["bwd"] * 1e3,
                                        "frac": bb / (res["bwd"] * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
            # the same kernels where HBM is the relevant bound: a large batch of reference station graphs
            res2, bf2, bb2, n2, e2 = measure_aggregation(dev, iters=10, which="ref4096")
            line["roofline"]["batched_reference_graphs"] = {
                "workload": f"4096 reference graphs in one batch: {n2} nodes, {e2} edges, H=128",
                "fwd": {"achieved": bf2 / (res2["fwd"] * 1e-3) / 1e9, "frac": bf2 / (res2["fwd"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "us_per_launch": res2["fwd"] * 1e3, "algorithmic_bytes": bf2},
                "bwd": {"achieved": bb2 / (res2["bwd"] * 1e-3) / 1e9, "frac": bb2 / (res2["bwd"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "us_per_launch": res2["bwd"] * 1e3, "algorithmic_bytes": bb2}}
            try:        # second north-star kernel: the DeepSets member contraction on tcgen05 (reported, not the headline)
                line["roofline"]["deepsets_contraction"] = measure_deepsets_contraction(dev, peaks)
            except Exception as exc:        # never lose the bench line over the extra leg
                line["roofline"]["deepsets_contraction"] = {"error": repr(exc)}
            try:        # third north-star kernel: links + mixture CRPS
                line["roofline"]["crps"] = measure_crps(dev, peaks)
            except Exception as exc:
                line["roofline"]["crps"] = {"error": repr(exc)}
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    tr = json.load(f)
                line["roofline"]["traffic"] = tr.get("gine_aggr_fwd_dram_bytes")
                line["roofline"]["bwd"]["traffic"] = tr.get("gine_aggr_bwd_dram_bytes")
                line["roofline"]["traffic_source"] = "profiles/traffic.json (one ncu --set full capture of the same kernels; not re-measured by this run)"
            except OSError:
                pass
        if not args.no_roofline:
            try:
                fl, by, n_par = step_flops_bytes()
                fp32_peak = measure_fp32_peak(dev)
                t_step = ms_dev / K_steps * 1e-3
                t_floor = max(fl / (fp32_peak * 1e12), by / (peaks["hbm_gbs"] * 1e9))
                line["step_roofline"] = {"flops_per_step": fl, "bytes_per_step": by, "achieved_tflops": fl / t_step / 1e12,
                                         "fp32_fma_peak_tflops_measured": fp32_peak, "frac_fp32_peak": fl / t_step / 1e12 / fp32_peak,
                                         "achieved_gbs": by / t_step / 1e9, "frac_hbm": by / t_step / 1e9 / peaks["hbm_gbs"],
                                         "floor_us": t_floor * 1e6, "frac_of_floor": t_floor / t_step,
                                         "note": "the reference-shape step (976 nodes) is bound by neither: it is a chain of ~45 dependent kernels "
                                                 "of a few microseconds (launch / dependency latency); see config4 for the throughput regime"}
            except Exception as exc:
                line["step_roofline"] = {"error": repr(exc)}
        if world == 1 and not args.no_roofline:
            for key, fn in (("gpu_eager_baseline", lambda: measure_gpu_eager(dev)),
                            ("config4", lambda: dict(measure_scaled_step(dev, HIDDEN, False), workload="one 100k-node graph, 2 978 560 edges, 51 members, H=128, L=4, fp32 (3xTF32 tensor cores)")),
                            ("config5", lambda: dict(measure_scaled_step(dev, 512, True), workload="config-4 graph, bf16 DeepSets (tcgen05 kind::f16), H=512, L=4")),
                            ("config5_reference_shape", lambda: dict(measure_refshape_step(dev, B_PER_GPU, 512, True, flush), workload="B=8 reference graphs x 11 members, bf16 DeepSets (tcgen05 kind::f16), H=512, L=4; one CUDA graph, L2 flushed between steps")),
                            ("train_loop_e2e", lambda: measure_train_loop(dev)),
                            ("train_loop_resident", lambda: measure_train_loop_resident(dev))):
                try:
                    line[key] = fn()
                except Exception as exc:            # never lose the bench line over an extra leg
                    line[key] = {"error": repr(exc)}
        if world == 1 and args.batch_sweep:
            try:
                line["batch_sweep"] = measure_batch_sweep(dev)
            except Exception as exc:
                line["batch_sweep"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            sps, done, dt, cores = cpu_steps_per_second(10_000, 3, budget_s=12.0)
            line["cpu_baseline"] = {"value": sps * B_PER_GPU, "unit": "graphs/s", "cores": cores, "kind": "port",
                                    "sample": f"{done} train steps of B=8 graphs in {dt:.1f} s (oracle port, torch CPU, {cores} threads)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
