#!/usr/bin/env python3
"""Caller of the inference path with eval.py's command line and outputs (eval.py:30-40,75-237): every checkpoint in
<dir>/models is loaded strictly, predictions are averaged over checkpoints, the CRPS of the average is written to
<dir>/results/<data>_results.txt and the predictions to <dir>/results/<data>.csv.

    python -m raincast_gnn_b200.eval --leadtime 24h --dir trained_models/24h_mixed_u --data rf [--synthetic 16]
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys

import pandas as pd
import torch
from torch.optim import AdamW

from .models.gnn import GNN
from .pyg_compat import DataLoader
from .utils.data import split_graph
from .utils.dataset import EUPPBench, SyntheticEUPPBench


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Evaluate trained checkpoints (B200 kernels behind the reference API).")
    p.add_argument("--leadtime", type=str, default="24h")
    p.add_argument("--dir", type=str, required=True)
    p.add_argument("--data", type=str, default="rf", choices=["rf", "f"])
    p.add_argument("--batch_size_rf", type=int, default=8)
    p.add_argument("--batch_size_f", type=int, default=5)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--root_raw", type=str, default="data/EUPPBench/raw")
    p.add_argument("--root_processed", type=str, default="data/EUPPBench/processed")
    p.add_argument("--synthetic", type=int, default=0)
    return p.parse_args(argv)


@torch.no_grad()
def predict_model(model, loader, device):
    """eval.py:57-69."""
    model.eval()
    return torch.cat([model(batch.to(device)).cpu() for batch in loader], dim=0)


def main(argv=None):
    args = parse_args(argv)
    os.makedirs(os.path.join(args.dir, "logs"), exist_ok=True)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s [%(levelname)s] %(message)s", force=True,
                        handlers=[logging.StreamHandler(sys.stdout),
                                  logging.FileHandler(os.path.join(args.dir, "logs", f"eval_{args.data}.log"), mode="w")])
    logger = logging.getLogger(__name__)
    torch.manual_seed(args.seed)
    with open(os.path.join(args.dir, "params.json")) as f:
        cfg = json.load(f)
    if not torch.cuda.is_available():
        logger.error("A CUDA device is required: this implementation has no CPU path.")
        sys.exit(1)
    device = torch.device("cuda")
    if args.synthetic > 0:
        dataset = SyntheticEUPPBench(n_dates=args.synthetic, members=51 if args.data == "f" else 11,
                                     max_dist=cfg.get("max_dist", 100.0), seed=args.seed + 1)
    else:
        dataset = EUPPBench(root_raw=args.root_raw, root_processed=args.root_processed, leadtime=args.leadtime,
                            max_dist=cfg.get("max_dist", 100.0), split="test_rf" if args.data == "rf" else "test_f")
    graphs = [dataset[i] for i in range(len(dataset))]
    if args.data == "f":                          # eval.py:130-137: 51 members -> 5 graphs of 10
        graphs = [g for d in graphs for g in split_graph(d, True)]
    loader = DataLoader(graphs, batch_size=args.batch_size_rf if args.data == "rf" else args.batch_size_f, shuffle=False)
    targets = torch.cat([d.y for d in graphs], dim=0)
    ckpt_dir = os.path.join(args.dir, "models")
    files = sorted(f for f in os.listdir(ckpt_dir) if f.endswith(".ckpt") or f.endswith(".pth")) if os.path.isdir(ckpt_dir) else []
    if not files:
        logger.error("No checkpoints found in %s", ckpt_dir)
        sys.exit(1)
    preds_all = []
    for name in files:
        model = GNN(in_channels=graphs[0].x.shape[1], hidden_channels_gnn=cfg["gnn_hidden"], out_channels_gnn=cfg["gnn_hidden"],
                    num_layers_gnn=cfg["gnn_layers"], optimizer_class=AdamW, optimizer_params={"lr": cfg["lr"]},
                    loss=cfg["loss"], grad_u=cfg["grad_u"], u=cfg["u"], xi=cfg["xi"]).to(device)
        model.load_state_dict(torch.load(os.path.join(ckpt_dir, name), map_location=device))     # strict, eval.py:196-197
        preds_all.append(predict_model(model, loader, device))
    final = torch.stack(preds_all).mean(dim=0) if len(preds_all) > 1 else preds_all[0]
    crps = model.loss_fn.crps(final, targets)          # host tensors: scored by the CUDA kernel after a copy (eval.py:213)
    logger.info(f"Final CRPS for data='{args.data}': {crps.item():.6f}")
    out_dir = os.path.join(args.dir, "results")
    os.makedirs(out_dir, exist_ok=True)
    pd.DataFrame(final.numpy()).to_csv(os.path.join(out_dir, f"{args.data}.csv"), index=False)
    with open(os.path.join(out_dir, f"{args.data}_results.txt"), "w") as f:
        f.write(f"CRPS: {crps.item():.6f}\n")
    return crps.item()


if __name__ == "__main__":
    main()
