"""Device time of the data-parallel gradient exchange + AdamW (rc_p2p_step, or NCCL all-reduce + rc_adamw_step with
RC_DP_EXCHANGE=nccl), measured with CUDA events around the optimiser call of eager steps.  Run under torchrun (>= 2 GPUs)."""
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
pg = dp.init_from_env("nccl")
ei, ea, ei_b, ea_b = B.static_graph(B.B_PER_GPU)
m = B.B_PER_GPU * B.N_STATIONS
eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), build_station_graph(ei_b, ea_b, m).to(dev), m, B.MEMBERS, B.FEATS, lr=1e-4,
                  process_group=pg, use_cuda_graph=False)
batches = B.make_host_batches(2, B.B_PER_GPU, seed=7, rank=rank, world=world)
eng.load_batch(*batches[0])
orig = eng._optimizer
times = []


def timed():
    torch.distributed.all_reduce(torch.zeros(1, device=dev), group=pg)     # align the ranks: measure the exchange, not the skew
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig()
    c.record()
    times.append((a, c))


eng._optimizer = timed
from raincast_gnn_b200 import _lib
trace = torch.zeros(8, dtype=torch.int64, device=dev)
for it in range(30):
    if it == 25:
        _lib.lib().rc_debug_p2p_trace(trace.data_ptr())
    eng.step()
torch.cuda.synchronize()
_lib.lib().rc_debug_p2p_trace(None)
t = trace.cpu().tolist()
if t[0]:
    print(f"rank {rank}: phases of the last exchange kernel (us after entry): published {1e-3*(t[1]-t[0]):.1f}, all arrived {1e-3*(t[2]-t[0]):.1f}, "
          f"update done {1e-3*(t[3]-t[0]):.1f}, exit {1e-3*(t[4]-t[0]):.1f}; entry at {t[0] % 10**9} ns", flush=True)
ms = sorted(a.elapsed_time(c) for a, c in times[10:])
print(f"rank {rank}: exchange + AdamW ({'peer memory' if eng.p2p is not None else 'nccl'}): median {ms[len(ms)//2]*1e3:.1f} us, min {ms[0]*1e3:.1f} us", flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
