#!/usr/bin/env python3
"""Training entry point with the reference's command line (train.py:30-38), run-directory layout
(<dir>/params.json, <dir>/logs/train_<run_id>.log, <dir>/models/run_<run_id>-best.ckpt) and epoch loop
(train.py:55-91,198-208), driving the B200 kernels.

    python -m raincast_gnn_b200.train --leadtime 24h --dir runs/24h_mixed_u --run_id 0 [--synthetic 64] [--host_loader | --autograd]

Additions (all optional): `--synthetic N` trains on N synthetic forecast dates of the reference shape (the
EUPPBench files need network access), the CUDA-graph engine steps the loop by default (same arithmetic, no
per-step host sync; under torchrun it shards the dates and exchanges the gradients) with the training split held on
the GPU (`--host_loader`: per-step host collate + H2D as in the reference; `--autograd`: the reference's loop verbatim),
`--max_epochs` overrides params.json.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import random
import sys

import numpy as np
import torch
from torch.utils.data import Subset, random_split

from . import dp
from .engine import TrainEngine
from .models.gnn import GNN
from .pyg_compat import DataLoader
from .utils.dataset import DeviceSplit, EUPPBench, SyntheticEUPPBench

LOG = logging.getLogger("raincast_gnn_b200.train")


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Train the DeepSets + GINE post-processing model on B200 kernels.")
    for flag, kw in (("--leadtime", dict(type=str, default="24h")),
                     ("--dir", dict(type=str, required=True, help="run directory holding params.json")),
                     ("--run_id", dict(type=str, required=True)),
                     ("--seed", dict(type=int, default=42)),
                     ("--root_raw", dict(type=str, default="data/EUPPBench/raw")),
                     ("--root_processed", dict(type=str, default="data/EUPPBench/processed")),
                     ("--synthetic", dict(type=int, default=0, help="number of synthetic dates (0: EUPPBench files)")),
                     ("--max_epochs", dict(type=int, default=None))):
        ap.add_argument(flag, **kw)
    ap.add_argument("--engine", action="store_true", default=True,
                    help="CUDA-graph training engine: the explicit kernel schedule of one train.py iteration, fused AdamW (default)")
    ap.add_argument("--autograd", dest="engine", action="store_false",
                    help="the reference's loop verbatim: module forward, loss.backward() through torch.autograd, torch.optim.AdamW")
    ap.add_argument("--resident", dest="resident", action="store_true", default=None,
                    help="with the engine: keep the training split on the GPU and build batches there (no per-step collate / H2D); "
                         "the default whenever the split fits in a quarter of the free device memory")
    ap.add_argument("--host_loader", dest="resident", action="store_false",
                    help="with the engine: collate every batch on the host and copy it to the GPU, as the reference's DataLoader does")
    return ap.parse_args(argv)


def seed_everything(seed: int):
    """train.py:44-49."""
    for fn in (random.seed, np.random.seed, torch.manual_seed):
        fn(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _open_log(run_dir: str, run_id: str, rank: int):
    os.makedirs(os.path.join(run_dir, "logs"), exist_ok=True)
    sinks = [logging.StreamHandler(sys.stdout)]
    if rank == 0:
        sinks.append(logging.FileHandler(os.path.join(run_dir, "logs", f"train_{run_id}.log"), mode="w"))
    logging.basicConfig(level=logging.INFO, format="%(asctime)s [%(levelname)s] %(message)s", handlers=sinks, force=True)


def run_epoch_autograd(model, loader, optimizer, device) -> float:
    """One pass of train.py:55-74; the loss is accumulated on the device and read once."""
    model.train()
    acc = torch.zeros((), dtype=torch.float64, device=device)
    for batch in loader:
        batch = batch.to(device)
        loss = model.loss_fn.crps(model(batch), batch.y)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        acc += loss.detach()
    return acc.item() / max(len(loader), 1)


def run_epoch_engine(engine: TrainEngine, loader) -> float:
    """The same pass through the captured step: the next batch's H2D copy of x / ensemble / y runs on a copy stream
    while the current step replays, then (all-reduce,) fused AdamW."""
    from .pyg_compat.data import Batch
    was_pinning, Batch.pin_outputs = Batch.pin_outputs, torch.cuda.is_available()   # collate x / ensemble / y straight into pinned memory
    try:
        return _run_epoch_engine(engine, loader)
    finally:
        Batch.pin_outputs = was_pinning


def _run_epoch_engine(engine: TrainEngine, loader) -> float:
    engine.loss_sum.zero_()
    done = 0
    pinned = torch.cuda.is_available()
    ragged = []                                            # the last batch of an epoch may be smaller (train.py:61-71 trains on it)

    def full_batches():
        for b in loader:
            if b.x.shape[0] == engine.m:
                yield b
            else:
                ragged.append(b)
    it = full_batches()

    def start(batch):
        host = [t.pin_memory() if pinned and t.device.type == "cpu" and not t.is_pinned() else t for t in (batch.x, batch.ensemble, batch.y)]
        engine.prefetch(*host)
        return host                                  # keeps the pinned tensors alive until the copy has run
    batch = next(it, None)
    held = start(batch) if batch is not None else None
    while batch is not None:
        engine.take_prefetched()
        batch = next(it, None)
        nxt = start(batch) if batch is not None else None
        engine.step()
        held = nxt
        done += 1
    del held
    for b in ragged:                                       # same kernels, stepped outside the captured graph (once per epoch)
        b = b.to(engine.device)
        engine.step_eager(b.x, b.ensemble, b.y, b.station_graph)
        done += 1
    return engine.loss_sum.item() / max(done, 1)


def run_epoch_resident(engine: TrainEngine, split: DeviceSplit, batch_size: int, generator=None) -> float:
    """One pass over a GPU-resident split: a device-side gather builds every batch.  With
    `generator=None` the epoch's order is drawn from torch's global RNG exactly as `pyg_compat.DataLoader(shuffle=True)`
    draws it, so this loop and `run_epoch_engine` step through the same batches."""
    engine.loss_sum.zero_()
    batches = split.epoch_batches(batch_size, generator=generator)
    full = [d for d in batches if int(d.numel()) * split.num_stations == engine.m]
    if full and engine.use_cuda_graph:
        # the epoch's order goes to the device once; every step is then ONE graph replay (gather of the next batch + the
        # training step): no host work, copy or kernel launch between the steps
        engine.begin_epoch(split, torch.stack(full))
        for _ in full:
            engine.step_resident()
    else:
        for dates in full:                                 # (an engine without CUDA graphs: one gather launch per step)
            engine.load_dates(split, dates)
            engine.step()
    for dates in batches:
        if int(dates.numel()) * split.num_stations != engine.m:          # ragged last batch
            b = int(dates.numel())
            g = split.batched_graph(b)
            engine.step_eager(split.x[dates].reshape(-1, split.x.shape[-1]), split.ensemble[dates].reshape(-1, *split.ensemble.shape[2:]),
                              split.y[dates].reshape(-1), g)
    mean_loss = engine.loss_sum.item() / max(len(batches), 1)
    engine.check_dates()
    return mean_loss


@torch.no_grad()
def validate(model, loader, device) -> float:
    """train.py:76-91 (BatchNorm on running statistics)."""
    model.eval()
    acc = torch.zeros((), dtype=torch.float64, device=device)
    for batch in loader:
        batch = batch.to(device)
        acc += model.loss_fn.crps(model(batch), batch.y)
    return acc.item() / max(len(loader), 1)


def main(argv=None):
    args = parse_args(argv)
    rank, local_rank, world = dp.env_world()
    _open_log(args.dir, args.run_id, rank)
    LOG.info("training run %s in %s: %s", args.run_id, args.dir, vars(args))
    seed_everything(args.seed)
    cfg_file = os.path.join(args.dir, "params.json")
    if not os.path.isfile(cfg_file):
        LOG.error("params.json is missing from %s", args.dir)
        sys.exit(1)
    with open(cfg_file) as fh:
        cfg = json.load(fh)
    LOG.info("params.json: %s", cfg)
    if not torch.cuda.is_available():
        LOG.error("no CUDA device: this implementation has no CPU path")
        sys.exit(1)
    group = dp.init_from_env("nccl")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)

    max_dist = cfg.get("max_dist", 100.0)
    if args.synthetic > 0:
        full = SyntheticEUPPBench(n_dates=args.synthetic, max_dist=max_dist, seed=args.seed)
    else:
        full = EUPPBench(root_raw=args.root_raw, root_processed=args.root_processed, leadtime=args.leadtime,
                         max_dist=max_dist, split="train_rf")
    held_out = int(0.1 * len(full))                                  # 90/10 split, train.py:149-153
    fit_part, val_part = random_split(full, [len(full) - held_out, held_out])
    if world > 1:                                                    # dates rank::world, equal counts per rank
        fit_part = Subset(fit_part, dp.shard_dates(len(fit_part), rank, world, seed=args.seed))
    LOG.info("dates: %d to fit, %d held out", len(fit_part), len(val_part))
    bs = cfg["batch_size"]
    fit_loader, val_loader = DataLoader(fit_part, batch_size=bs, shuffle=True), DataLoader(val_part, batch_size=bs)

    probe = fit_part[0]
    model = GNN(in_channels=probe.x.shape[1], hidden_channels_gnn=cfg["gnn_hidden"], out_channels_gnn=cfg["gnn_hidden"],
                num_layers_gnn=cfg["gnn_layers"], optimizer_class=torch.optim.AdamW, optimizer_params={"lr": cfg["lr"]},
                loss=cfg["loss"], grad_u=cfg["grad_u"], u=cfg["u"], xi=cfg["xi"]).to(device)
    with torch.no_grad():
        model(probe.to(device))          # the reference's pre-training forward on one un-batched graph (train.py:182-183)
    engine = optimizer = None
    if args.engine:
        first = next(iter(fit_loader))
        engine = TrainEngine(model, first.station_graph, first.x.shape[0], first.ensemble.shape[1], first.x.shape[1],
                             lr=cfg["lr"], process_group=group).capture()
    elif world > 1:
        raise SystemExit("data-parallel training needs the engine (the gradient exchange lives there): drop --autograd")
    else:
        optimizer = model.optimizer_class(model.parameters(), **model.optimizer_params)

    resident = None
    if args.resident and engine is None:
        raise SystemExit("--resident needs the engine: drop --autograd")
    if engine is not None and args.resident is not False:
        # 180 GB of HBM hold any split of this dataset (3.1k dates x 122 stations x 11 members x 35 features = 0.6 GB):
        # the batches are then gathered on the device, the host only draws the epoch's order
        split_bytes = 4 * len(fit_part) * (probe.x.numel() + probe.ensemble.numel() + probe.y.numel())
        if args.resident or split_bytes < torch.cuda.mem_get_info(device)[0] // 4:
            try:
                resident = DeviceSplit([fit_part[i] for i in range(len(fit_part))], device)
                LOG.info("training split resident on %s: %.1f MB", device, split_bytes / 1e6)
            except ValueError as exc:                                # graphs of several shapes: no static station graph
                if args.resident:
                    raise
                LOG.info("host loader (%s)", exc)
    os.makedirs(os.path.join(args.dir, "models"), exist_ok=True)
    target = os.path.join(args.dir, "models", f"run_{args.run_id}-best.ckpt")
    best, saved = float("inf"), None
    n_epochs = args.max_epochs or cfg["max_epochs"]
    for epoch in range(1, n_epochs + 1):
        if resident is not None:
            fit_loss = run_epoch_resident(engine, resident, bs)
        elif engine is not None:
            fit_loss = run_epoch_engine(engine, fit_loader)
        else:
            fit_loss = run_epoch_autograd(model, fit_loader, optimizer, device)
        val_loss = validate(model, val_loader, device) if len(val_part) else float("nan")
        LOG.info("epoch %d/%d  [Train] Loss: %.6f  [Val] Loss: %.6f", epoch, n_epochs, fit_loss, val_loss)
        if rank == 0 and (val_loss < best or saved is None):
            best, saved = val_loss, target
            torch.save(model.state_dict(), target)                   # bare state_dict, train.py:207
            LOG.info("[Checkpoint] val %.6f -> %s", val_loss, target)
    LOG.info("done; best checkpoint: %s", saved)
    return saved


if __name__ == "__main__":
    main()
