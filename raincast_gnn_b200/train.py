#!/usr/bin/env python3
"""Caller of the hot path with train.py's command line, logging and checkpoint layout (train.py:30-38,97-216).

    python -m raincast_gnn_b200.train --leadtime 24h --dir trained_models/24h_mixed_u --run_id 0 [--synthetic 64]

Identical flags to the reference; `--synthetic N` (new, optional) trains on N synthetic forecast dates of the
reference shape instead of the EUPPBench files (which need network access to obtain).  `--engine` (new,
optional) swaps the autograd loop for the CUDA-graph engine (same arithmetic, no per-step host sync); under
torchrun it shards the dates over the ranks and all-reduces the gradients (DDP semantics).
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
from torch.optim import AdamW
from torch.utils.data import random_split

from . import dp
from .engine import TrainEngine
from .models.gnn import GNN
from .pyg_compat import DataLoader
from .utils.dataset import EUPPBench, SyntheticEUPPBench


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Train a graph-based model (B200 kernels behind the reference API).")
    p.add_argument("--leadtime", type=str, default="24h")
    p.add_argument("--dir", type=str, required=True, help="Directory containing params.json and for logs/checkpoints.")
    p.add_argument("--run_id", type=str, required=True)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--root_raw", type=str, default="data/EUPPBench/raw")
    p.add_argument("--root_processed", type=str, default="data/EUPPBench/processed")
    p.add_argument("--synthetic", type=int, default=0, help="train on this many synthetic dates (no dataset files needed)")
    p.add_argument("--engine", action="store_true", help="use the CUDA-graph training engine")
    p.add_argument("--max_epochs", type=int, default=None, help="override params.json max_epochs")
    return p.parse_args(argv)


def set_seed(seed: int):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def train_one_epoch(model, loader, optimizer, device, logger):
    """train.py:55-74, with the per-step `.item()` sync replaced by one device-side accumulator read per epoch."""
    model.train()
    total = torch.zeros((), dtype=torch.float64, device=device)
    for batch in loader:
        batch = batch.to(device)
        preds = model(batch)
        loss = model.loss_fn.crps(preds, batch.y)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        total += loss.detach()
    avg = total.item() / max(len(loader), 1)
    logger.info(f"  [Train] Loss: {avg:.6f}")
    return avg


def train_one_epoch_engine(engine, loader, logger):
    """Same epoch through the graphed engine: H2D of x / ensemble / y, graph replay, all-reduce, fused AdamW."""
    engine.loss_sum.zero_()
    steps = 0
    for batch in loader:
        if batch.x.shape[0] != engine.m:          # ragged last batch: the captured graph has a fixed shape
            continue
        engine.load_batch(batch.x, batch.ensemble, batch.y)
        engine.step()
        steps += 1
    avg = engine.loss_sum.item() / max(steps, 1)
    logger.info(f"  [Train] Loss: {avg:.6f}")
    return avg


def evaluate(model, loader, device, logger):
    """train.py:76-91."""
    model.eval()
    total = torch.zeros((), dtype=torch.float64, device=device)
    with torch.no_grad():
        for batch in loader:
            batch = batch.to(device)
            total += model.loss_fn.crps(model(batch), batch.y)
    avg = total.item() / max(len(loader), 1)
    logger.info(f"  [Val] Loss: {avg:.6f}")
    return avg


def main(argv=None):
    args = parse_args(argv)
    rank, local_rank, world = dp.env_world()
    os.makedirs(os.path.join(args.dir, "logs"), exist_ok=True)
    handlers = [logging.StreamHandler(sys.stdout)]
    if rank == 0:
        handlers.append(logging.FileHandler(os.path.join(args.dir, "logs", f"train_{args.run_id}.log"), mode="w"))
    logging.basicConfig(level=logging.INFO, format="%(asctime)s [%(levelname)s] %(message)s", handlers=handlers, force=True)
    logger = logging.getLogger(__name__)
    logger.info("========== Training Script Started ==========")
    logger.info(f"Arguments: {args}")
    set_seed(args.seed)
    config_path = os.path.join(args.dir, "params.json")
    if not os.path.isfile(config_path):
        logger.error(f"Could not find params.json at: {config_path}")
        sys.exit(1)
    with open(config_path) as f:
        config = json.load(f)
    logger.info(f"Loaded config: {config}")
    if not torch.cuda.is_available():
        logger.error("A CUDA device is required: this implementation has no CPU path.")
        sys.exit(1)
    group = dp.init_from_env("nccl")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)

    if args.synthetic > 0:
        dataset = SyntheticEUPPBench(n_dates=args.synthetic, max_dist=config.get("max_dist", 100.0), seed=args.seed)
    else:
        dataset = EUPPBench(root_raw=args.root_raw, root_processed=args.root_processed, leadtime=args.leadtime,
                            max_dist=config.get("max_dist", 100.0), split="train_rf")
    n_total = len(dataset)
    n_val = int(0.1 * n_total)
    train_set, val_set = random_split(dataset, [n_total - n_val, n_val])
    if world > 1:                                  # forecast dates sharded rank::world, same count on every rank
        mine = dp.shard_dates(len(train_set), rank, world, seed=args.seed)
        train_set = torch.utils.data.Subset(train_set, mine)
    logger.info(f"Dataset sizes => Train: {len(train_set)}, Val: {len(val_set)}")
    train_loader = DataLoader(train_set, batch_size=config["batch_size"], shuffle=True)
    val_loader = DataLoader(val_set, batch_size=config["batch_size"], shuffle=False)

    example = train_set[0]
    model = GNN(in_channels=example.x.shape[1], hidden_channels_gnn=config["gnn_hidden"], out_channels_gnn=config["gnn_hidden"],
                num_layers_gnn=config["gnn_layers"], optimizer_class=AdamW, optimizer_params={"lr": config["lr"]},
                loss=config["loss"], grad_u=config["grad_u"], u=config["u"], xi=config["xi"]).to(device)
    with torch.no_grad():                          # train.py:182-183: one forward on a single un-batched graph
        model(example.to(device))
    max_epochs = args.max_epochs or config["max_epochs"]
    engine = optimizer = None
    if args.engine:
        first = next(iter(train_loader))
        engine = TrainEngine(model, first.station_graph, first.x.shape[0], first.ensemble.shape[1], first.x.shape[1],
                             lr=config["lr"], process_group=group).capture()
    else:
        if world > 1:
            raise SystemExit("data-parallel training needs --engine (the gradient all-reduce lives in the engine)")
        optimizer = model.optimizer_class(model.parameters(), **model.optimizer_params)

    ckpt_dir = os.path.join(args.dir, "models")
    os.makedirs(ckpt_dir, exist_ok=True)
    best_val, best_path = float("inf"), None
    logger.info(f"Starting training for {max_epochs} epochs...")
    for epoch in range(1, max_epochs + 1):
        logger.info(f"=== Epoch {epoch}/{max_epochs} ===")
        if engine is not None:
            train_one_epoch_engine(engine, train_loader, logger)
        else:
            train_one_epoch(model, train_loader, optimizer, device, logger)
        val_loss = evaluate(model, val_loader, device, logger) if len(val_set) else float("nan")
        if rank == 0 and (val_loss < best_val or best_path is None):
            best_val = val_loss
            best_path = os.path.join(ckpt_dir, f"run_{args.run_id}-best.ckpt")
            torch.save(model.state_dict(), best_path)          # bare state_dict, train.py:207
            logger.info(f"[Checkpoint] New best val_loss: {val_loss:.6f}. Saved to {best_path}")
    logger.info("Training completed.")
    logger.info("========== Training Script Finished ==========")
    return best_path


if __name__ == "__main__":
    main()
