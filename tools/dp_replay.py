"""Back-to-back graph replays of the data-parallel reference-shape step (no L2 flush, no re-alignment): device time per step
on every rank.  torchrun --nproc-per-node N tools/dp_replay.py"""
import os, sys
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
pg = dp.init_from_env("nccl") if world > 1 else None
if world == 1 and os.environ.get("RC_DP_FORCE_P2P") == "1":          # the exchange path with one rank (measurement)
    torch.distributed.init_process_group(backend="nccl", rank=0, world_size=1, device_id=dev)
    pg = torch.distributed.group.WORLD
ei, ea, ei_b, ea_b = B.static_graph(B.B_PER_GPU)
m = B.B_PER_GPU * B.N_STATIONS
from raincast_gnn_b200 import _lib
trace = torch.zeros(8, dtype=torch.int64, device=dev)

if pg is not None and os.environ.get("RC_TRACE", "0") == "1":
    _lib.lib().rc_debug_p2p_trace(trace.data_ptr())       # before capture: the pointer is a kernel argument of the graph
eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), build_station_graph(ei_b, ea_b, m).to(dev), m, B.MEMBERS, B.FEATS, lr=1e-4,
                  process_group=pg).capture()
batches = B.make_host_batches(2, B.B_PER_GPU, seed=7, rank=rank, world=world)
eng.load_batch(*batches[0])
for _ in range(20):
    eng.step()
torch.cuda.synchronize()
if pg is not None:
    torch.distributed.barrier()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(300):
    eng.step()
c.record(); c.synchronize()
print(f"rank {rank}/{world}: {a.elapsed_time(c) / 300 * 1e3:.1f} us per step back to back ({'p2p' if eng.p2p is not None else ('nccl' if world > 1 else 'single')}), launches {eng.kernels_per_step}", flush=True)
if pg is not None and os.environ.get("RC_TRACE", "0") == "1":
    # two more steps, reading the stamps after each: step start (wait_done entry) -> exchange entry -> exit -> next start
    t2 = trace.cpu().tolist(); t1 = t2       # stamps of the LAST step of the back-to-back run (steady state)
    us = lambda a, b: (b - a) * 1e-3
    print(f"rank {rank}: step start -> wait_done exit {us(t2[5], t2[6]):.1f}; start -> exchange entry {us(t2[5], t2[0]):.1f}; exchange: published {us(t2[0], t2[1]):.1f}, "
          f"all arrived {us(t2[0], t2[2]):.1f}, update done {us(t2[0], t2[3]):.1f}, exit {us(t2[0], t2[4]):.1f}; whole step (start -> exchange exit) {us(t2[5], t2[4]):.1f}", flush=True)
if pg is not None:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
