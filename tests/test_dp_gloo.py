"""N > 1 host logic on the CPU with the gloo backend (world_size 2): date sharding and the DDP semantics of the
gradient all-reduce (mean of per-rank gradients, per-rank BatchNorm statistics and per-rank valid-node means).
The per-rank gradients come from the CPU oracle here; the CUDA engine plugs the same flat buffer into the same call."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raincast_gnn_b200 import dp


def test_shard_dates_partitions_every_epoch():
    for world in (1, 2, 4, 8):
        for epoch in (0, 1):
            parts = [dp.shard_dates(103, r, world, seed=3, epoch=epoch) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert len(set(flat)) == len(flat) == (103 // world) * world
            assert len({len(p) for p in parts}) == 1
    assert dp.shard_dates(50, 0, 2, seed=3, epoch=0) != dp.shard_dates(50, 0, 2, seed=3, epoch=1)
    assert dp.shard_dates(50, 1, 2, seed=3, epoch=0) == dp.shard_dates(50, 1, 2, seed=3, epoch=0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tiny_setup():
    from oracle import graph as og, model as om
    from raincast_gnn_b200.utils import synthetic as syn
    n, em, f, h = 12, 3, 5, 16
    ei, ea = og.radius_graph(syn.distance_matrix(syn.station_coords(n, 600.0, 0)), 300.0)
    kw = dict(in_channels=f, hidden_channels_gnn=h, out_channels_gnn=h, num_layers_gnn=2, loss="MixedLoss", grad_u="True",
              u=1.71, xi=0.5)
    model = om.GNN(**kw)
    model.load_state_dict(syn.seeded_state_dict(model.state_dict(), seed=11))
    return n, em, f, torch.from_numpy(ei), torch.from_numpy(ea), model


def _rank_gradient(model, dates, n, em, f, ei, ea):
    from oracle import pyg as opyg
    from raincast_gnn_b200.utils import synthetic as syn
    items = []
    for d in dates:
        x, ens = syn.node_features(n, em, f, seed=100 + d)
        items.append(opyg.Data(x=x, ensemble=ens, edge_index=ei, edge_attr=ea, y=syn.log_precip_targets(n, seed=100 + d)))
    batch = opyg.Batch.from_data_list(items)
    model.zero_grad()
    model.train()
    loss = model.loss_fn.crps(model(batch), batch.y)
    loss.backward()
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()])


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    group = dp.init_from_env("gloo")
    assert dp.env_world() == (rank, rank, world)
    n, em, f, ei, ea, model = _tiny_setup()
    dates = dp.shard_dates(8, rank, world, seed=1)
    flat = _rank_gradient(model, dates, n, em, f, ei, ea)
    scale = dp.allreduce_mean_(flat, group, fold_scale=True)       # the AdamW kernel folds the 1/world factor
    np.save(os.path.join(out_dir, f"g{rank}.npy"), (flat * scale).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_has_ddp_semantics(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"g{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1])
    torch.set_num_threads(1)
    n, em, f, ei, ea, model = _tiny_setup()
    want = sum(_rank_gradient(model, dp.shard_dates(8, r, world, seed=1), n, em, f, ei, ea) for r in range(world)) / world
    assert np.abs(got[0] - want.numpy()).max() <= 1e-6 * np.abs(want.numpy()).max()
