"""Data-parallel step on real GPUs (needs >= 2; the CPU / gloo coverage of the sharding logic is tests/test_dp_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("strict_flags", ["0", "1"])
def test_peer_memory_exchange_matches_nccl(strict_flags):
    """rc_p2p_step (gradients summed straight from the peers' memory, in rank order, inside the AdamW kernel) against one
    NCCL all-reduce + rc_adamw_step: same loss trajectories to 1e-5, bit-identical replicas on every rank - with the
    default flag protocol (device-scope fences, rc_p2p_flag_scope(1)) and with release / acquire at system scope."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, NCCL_DEBUG_FILE="/dev/stderr", RC_P2P_STRICT=strict_flags)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")], capture_output=True, text=True,
                         timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["p2p_replicas_identical"] and out["nccl_replicas_identical"]
    assert out["loss_rel_diff"] < 1e-5 and out["steps"] == 6
    assert out["launches_per_step"][0] == out["launches_per_step"][1]          # one exchange + AdamW kernel  vs  AdamW after the all-reduce


@pytest.mark.gpu
def test_data_parallel_semantics_on_one_gpu_vs_micro_batch_oracle():
    """DDP semantics without a second GPU: G ranks are G micro-batches - per-micro-batch BatchNorm
    statistics, per-micro-batch mean over valid nodes, MEAN of the G gradients, one AdamW step.  The engine's emulation
    (the arithmetic of rc_p2p_step: gradients summed in rank order, grad_scale = 1/G inside AdamW) against the oracle:
    G independent forward / backward passes of the reference arithmetic from the same weights, gradients averaged,
    torch.optim.AdamW.  Losses, and every parameter after the step, to 1e-5 (of the parameter's scale; Adam's first step
    moves every weight by ~lr * sign(grad), so a parameter is compared on max(|p|, lr))."""
    import numpy as np
    import torch
    from oracle import model as om, pyg as opyg
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import DataLoader
    from raincast_gnn_b200.utils import synthetic as syn
    from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda:0")
    G, B = 4, 2
    ds = SyntheticEUPPBench(n_dates=G * B, num_stations=40, members=11, feats=9, max_dist=200.0)
    loader = list(DataLoader(ds, batch_size=B))
    kw = dict(in_channels=9, hidden_channels_gnn=128, out_channels_gnn=128, num_layers_gnn=2, optimizer_class=torch.optim.AdamW,
              optimizer_params={"lr": 1e-3}, loss="MixedLoss", grad_u="True", u=1.71, xi=0.5)
    ours = GNN(**kw)
    sd = syn.seeded_state_dict(ours.state_dict(), seed=7)
    ours.load_state_dict(sd)
    ours.to(dev).train()
    eng = TrainEngine(ours, loader[0].station_graph, loader[0].x.shape[0], 11, 9, lr=1e-3, use_cuda_graph=False)
    losses, mean_grads = eng.step_emulated_ranks([(b.x, b.ensemble, b.y) for b in loader])
    losses = losses.cpu()
    # oracle: G ranks = G independent passes from the same weights, mean of the gradients, AdamW
    ref = om.GNN(**kw).double()
    ref.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()})
    ref.conv.force_float = False
    ref.train()
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    acc = {k: torch.zeros_like(p) for k, p in ref.named_parameters()}
    ref_losses = []
    for b in loader:
        d = opyg.Data(x=b.x.double(), ensemble=b.ensemble.double(), edge_index=b.edge_index, edge_attr=b.edge_attr.double(), y=b.y.double())
        opt.zero_grad()
        l = ref.loss_fn.crps(ref(d), d.y)
        l.backward()
        ref_losses.append(float(l))
        for k, p in ref.named_parameters():
            acc[k] += p.grad
    named = dict(ref.named_parameters())
    for k, p in named.items():
        p.grad = acc[k] / G
    assert np.allclose(losses.numpy(), np.array(ref_losses), rtol=1e-5, atol=0)
    for k, p in named.items():                          # the mean gradient of the G ranks
        scale = named[k[:-4] + "weight"].grad.abs().max().item() if k.endswith(".nn.0.bias") else p.grad.abs().max().item()
        assert (mean_grads[k].cpu().double() - p.grad).abs().max().item() < 1e-5 * max(scale, 1e-12), k
    opt.step()
    for k, p in named.items():                          # and the AdamW step it feeds (first step: ~lr * sign(grad) per entry;
        if k.endswith(".nn.0.bias"):                    #  entries whose gradient is ~eps are ill-conditioned; the bias in front of
            continue                                    #  BatchNorm has a zero true gradient: Adam normalises pure rounding noise)
        got = dict(ours.named_parameters())[k].detach().cpu().double()
        firm = p.grad.abs() > 1e-4 * p.grad.abs().max()
        assert ((got - p.detach()).abs() * firm).max().item() < 1e-5 * max(p.detach().abs().max().item(), 1e-3), k
