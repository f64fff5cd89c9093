"""Data-parallel step on real GPUs (needs >= 2; the CPU / gloo coverage of the sharding logic is tests/test_dp_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_memory_exchange_matches_nccl():
    """rc_p2p_barrier + rc_p2p_adamw_step (gradients summed straight from the peers' memory, in rank order) against one
    NCCL all-reduce + rc_adamw_step: same loss trajectories to 1e-5, bit-identical replicas on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, NCCL_DEBUG_FILE="/dev/stderr")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")], capture_output=True, text=True,
                         timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["p2p_replicas_identical"] and out["nccl_replicas_identical"]
    assert out["loss_rel_diff"] < 1e-5 and out["steps"] == 6
    assert out["launches_per_step"][0] == out["launches_per_step"][1] + 2
