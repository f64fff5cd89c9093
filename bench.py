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
    ap.add_argument("--batch-sweep", action="store_true", help="also report graphs/s at B in {64, 512} per GPU")
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
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
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


def static_graph(batch: int):
    from raincast_gnn_b200 import graph as G
    from raincast_gnn_b200.utils import synthetic as syn
    ei, ea = G.radius_graph(syn.distance_matrix(syn.station_coords(N_STATIONS, 600.0, seed=0)), 100.0)
    assert ei.shape[1] == 1164, "synthetic reference graph must have 1 164 edges (SURVEY.md 8d)"
    ei_b, ea_b = G.collate_static(ei, ea, N_STATIONS, batch)
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
    ei, ea, ei_b, ea_b = static_graph(B_PER_GPU)
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
                                       ("summed from NVLink peer memory inside the AdamW kernel)" if eng.p2p is not None else "one NCCL all-reduce)"))
                                      if world > 1 else "single GPU (dates would be sharded rank::world; no gradient exchange)",
                       "timing": "per-step CUDA events on the launch stream; 256 MiB L2 flush between timed steps, outside the events; max over ranks",
                       "cuda_graph": True, "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / K_steps},
            "gpu_launches": eng.launches_per_step * K_steps, "launches_per_step": eng.launches_per_step,
            "clocks": clocks.summary(), "wall_s_device_loop": wall_dev}

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
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    tr = json.load(f)
                line["roofline"]["traffic"] = tr.get("gine_aggr_fwd_dram_bytes")
                line["roofline"]["bwd"]["traffic"] = tr.get("gine_aggr_bwd_dram_bytes")
            except OSError:
                pass
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
