"""Data-parallel plumbing: one process per GPU, forecast dates sharded across ranks, one all-reduce of the
flat gradient per step.  The reference has no distributed code; semantics are DDP's
(per-rank BatchNorm statistics, per-rank mean over valid nodes, mean of the rank gradients)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process: (0, 0, 1))."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* when WORLD_SIZE > 1; returns the group or None."""
    rank, local_rank, world = env_world()
    if world <= 1:
        return None
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return dist.group.WORLD


def shard_dates(n_dates: int, rank: int, world: int, seed: int = 0, epoch: int = 0, drop_last: bool = True):
    """Indices of the forecast dates rank `rank` trains on in `epoch`: `rank::world` of a seeded permutation
    (same on every rank).  With drop_last every rank gets the same count, so the ranks stay in lock step."""
    g = torch.Generator().manual_seed(seed * 1_000_003 + epoch)
    perm = torch.randperm(n_dates, generator=g)
    if drop_last:
        perm = perm[: (n_dates // world) * world]
    return perm[rank::world].tolist()


def allreduce_mean_(flat: torch.Tensor, group=None, fold_scale: bool = False) -> float:
    """Sum-all-reduce `flat` in place.  Returns the factor that turns the sum into the mean of the rank
    gradients; it is applied here unless fold_scale (the AdamW kernel folds it into its gradient read)."""
    world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    scale = 1.0 / world
    if not fold_scale and world > 1:
        flat.mul_(scale)
    return scale
