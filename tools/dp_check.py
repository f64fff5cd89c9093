"""Data-parallel check, run under torchrun with >= 2 GPUs (tests/test_gpu_dp.py starts it):
the peer-memory gradient exchange (rc_p2p_*) against one NCCL all-reduce + rc_adamw_step on the same batches.
Prints one JSON line on rank 0."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from raincast_gnn_b200 import dp
from raincast_gnn_b200.engine import TrainEngine
from raincast_gnn_b200.graph import build_station_graph
from raincast_gnn_b200.models import GNN

rank, local_rank, world = dp.env_world()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
pg = dp.init_from_env("nccl")
ei, ea, ei_b, ea_b = B.static_graph(B.B_PER_GPU)
m = B.B_PER_GPU * B.N_STATIONS
sg = build_station_graph(ei_b, ea_b, m).to(dev)
batches = B.make_host_batches(3, B.B_PER_GPU, seed=11, rank=rank, world=world)
steps = 6
out = {}
for mode in ("p2p", "nccl"):
    os.environ["RC_DP_EXCHANGE"] = mode
    model = B.seeded_model(GNN).to(dev).train()
    eng = TrainEngine(model, sg, m, B.MEMBERS, B.FEATS, lr=1e-3, process_group=pg).capture()
    assert (eng.p2p is not None) == (mode == "p2p"), f"exchange mode {mode} not in effect"
    losses = []
    for i in range(steps):
        eng.load_batch(*batches[i % len(batches)])
        losses.append(eng.step().clone())
    torch.cuda.synchronize()
    if mode == "p2p":
        eng.check_peers()
    # replicas must be bit-identical: gather every rank's parameters
    flat = eng.flat_p.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, flat, group=pg)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    loss_t = torch.stack(losses).reshape(-1)
    all_losses = [torch.empty_like(loss_t) for _ in range(world)]
    torch.distributed.all_gather(all_losses, loss_t, group=pg)
    out[mode] = {"replicas_identical": bool(same), "losses": [l.tolist() for l in all_losses], "params": flat.cpu(),
                 "step_count": int(eng.step_count), "launches": eng.launches_per_step}
    del eng, model
if rank == 0:
    lp, ln = torch.tensor(out["p2p"]["losses"]), torch.tensor(out["nccl"]["losses"])
    rel = float((lp - ln).abs().max() / ln.abs().max())
    dparam = float((out["p2p"]["params"] - out["nccl"]["params"]).abs().max())
    print(json.dumps({"world": world, "p2p_replicas_identical": out["p2p"]["replicas_identical"],
                      "nccl_replicas_identical": out["nccl"]["replicas_identical"], "loss_rel_diff": rel,
                      "param_max_abs_diff": dparam, "steps": out["p2p"]["step_count"],
                      "launches_per_step": [out["p2p"]["launches"], out["nccl"]["launches"]]}), flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
