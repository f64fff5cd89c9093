"""The reference's entry points (train.py / eval.py command lines, params.json, logs, .ckpt layout) driven end to
end on synthetic data, through the autograd module path and through the CUDA-graph engine."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def run_dir(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from raincast_gnn_b200.utils.configs import write_params_json
    d = tmp_path / "24h_mixed_u"
    write_params_json(str(d), "24h", "mixed_u")
    return str(d)


@pytest.mark.parametrize("engine", [False, True])
def test_train_then_eval_cli(run_dir, engine):
    from raincast_gnn_b200 import eval as rc_eval, train as rc_train
    argv = ["--leadtime", "24h", "--dir", run_dir, "--run_id", "0", "--synthetic", "40", "--max_epochs", "2"]
    ckpt = rc_train.main(argv + (["--engine"] if engine else ["--autograd"]))
    assert ckpt == os.path.join(run_dir, "models", "run_0-best.ckpt") and os.path.isfile(ckpt)
    log = open(os.path.join(run_dir, "logs", "train_0.log")).read()
    assert "[Train] Loss:" in log and "[Val] Loss:" in log and "[Checkpoint]" in log
    sd = torch.load(ckpt, map_location="cpu")
    assert len(sd) == 60 and sd["aggr.weight"].shape == (5, 128)
    assert int(sd["conv.convolutions.0.nn.1.num_batches_tracked"]) > 1
    for data in ("rf", "f"):
        crps = rc_eval.main(["--leadtime", "24h", "--dir", run_dir, "--data", data, "--synthetic", "4"])
        assert crps == crps and 0.0 < crps < 10.0
        assert os.path.isfile(os.path.join(run_dir, "results", f"{data}.csv"))
        assert open(os.path.join(run_dir, "results", f"{data}_results.txt")).read().startswith("CRPS:")


def test_training_reduces_the_loss(run_dir):
    """30 engine steps on one batch: the CRPS must go down (optimiser wiring sanity, not a parity check)."""
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import DataLoader
    from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
    cfg = json.load(open(os.path.join(run_dir, "params.json")))
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    batch = next(iter(DataLoader(SyntheticEUPPBench(n_dates=8), batch_size=8)))
    model = GNN(35, cfg["gnn_hidden"], cfg["gnn_hidden"], cfg["gnn_layers"], torch.optim.AdamW, {"lr": 1e-3}, cfg["loss"],
                cfg["grad_u"], cfg["u"], cfg["xi"]).to(dev).train()
    eng = TrainEngine(model, batch.station_graph, batch.x.shape[0], 11, 35, lr=1e-3).capture()
    eng.load_batch(batch.x, batch.ensemble, batch.y)
    losses = [float(eng.step().item()) for _ in range(30)]
    assert losses[-1] < 0.8 * losses[0]


def test_resident_split_matches_host_batches(run_dir):
    """GPU-resident split: rc_gather_dates builds the same batch as PyG-style collation + H2D, so the
    loss trajectories are identical; an out-of-range date index is reported, not read."""
    from raincast_gnn_b200 import _lib
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import Batch
    from raincast_gnn_b200.utils.dataset import DeviceSplit, SyntheticEUPPBench
    cfg = json.load(open(os.path.join(run_dir, "params.json")))
    dev = torch.device("cuda:0")
    ds = SyntheticEUPPBench(n_dates=24)
    split = DeviceSplit([ds[i] for i in range(len(ds))], dev)
    order = [torch.tensor(o) for o in ([3, 17, 0, 9, 22, 5, 11, 8], [1, 2, 4, 6, 7, 10, 12, 13], [23, 21, 20, 19, 18, 16, 15, 14])]
    trajs = []
    for resident in (False, True):
        torch.manual_seed(0)
        model = GNN(35, cfg["gnn_hidden"], cfg["gnn_hidden"], cfg["gnn_layers"], torch.optim.AdamW, {"lr": 1e-3}, cfg["loss"],
                    cfg["grad_u"], cfg["u"], cfg["xi"]).to(dev).train()
        first = Batch.from_data_list([ds[int(i)] for i in order[0]])
        eng = TrainEngine(model, first.station_graph, first.x.shape[0], 11, 35, lr=1e-3).capture()
        traj = []
        for dates in order:
            if resident:
                eng.load_dates(split, dates.to(dev))
            else:
                b = Batch.from_data_list([ds[int(i)] for i in dates])
                eng.load_batch(b.x, b.ensemble, b.y)
            traj.append(float(eng.step().item()))
        trajs.append(traj)
    assert trajs[0] == trajs[1]
    # the epoch as ONE graph per step (gather of batch number step_count - base + the step): same trajectory again, two
    # epochs back to back (the second in another order), and a replay past the epoch's batches is reported, not read
    torch.manual_seed(0)
    model = GNN(35, cfg["gnn_hidden"], cfg["gnn_hidden"], cfg["gnn_layers"], torch.optim.AdamW, {"lr": 1e-3}, cfg["loss"],
                cfg["grad_u"], cfg["u"], cfg["xi"]).to(dev).train()
    eng = TrainEngine(model, first.station_graph, first.x.shape[0], 11, 35, lr=1e-3).capture()
    eng.begin_epoch(split, torch.stack(order))
    traj = [float(eng.step_resident().item()) for _ in order]
    assert traj == trajs[0]
    eng.begin_epoch(split, torch.stack(order[::-1][:2]))
    again = [float(eng.step_resident().item()) for _ in range(2)]
    eng.check_dates()
    assert again[0] != traj[0] and all(l == l for l in again)
    eng.begin_epoch(split, torch.stack(order[:1]))
    eng.step_resident()
    eng.step_resident()
    with pytest.raises(_lib.RcError, match="past the batches"):
        eng.check_dates()
    eng.load_dates(split, torch.tensor([0, 1, 2, 3, 4, 5, 6, 24], device=dev))
    torch.cuda.synchronize()
    assert int(eng._bad_date) == 1
    with pytest.raises(_lib.RcError):
        eng.load_dates(split, torch.tensor([0, 1], device=dev))


def test_train_cli_resident(run_dir):
    """The unchanged command line keeps the training split on the GPU (the engine's default); `--host_loader` collates on
    the host like the reference's DataLoader.  Both draw the epoch order from torch's global RNG the same way, so the
    logged losses of the two runs are the same lines."""
    import re
    from raincast_gnn_b200 import train as rc_train
    base = ["--leadtime", "24h", "--dir", run_dir, "--synthetic", "43", "--max_epochs", "2"]
    losses = []
    for run_id, extra in (("1", []), ("2", ["--host_loader"]), ("3", ["--resident"])):
        ckpt = rc_train.main(base + ["--run_id", run_id] + extra)
        assert os.path.isfile(ckpt)
        log = open(os.path.join(run_dir, "logs", f"train_{run_id}.log")).read()
        assert ("resident on" in log) == (extra != ["--host_loader"]) and "[Train] Loss:" in log
        losses.append(re.findall(r"\[Train\] Loss: (\S+)\s+\[Val\] Loss: (\S+)", log))
    assert len(losses[0]) == 2 and losses[0] == losses[1] == losses[2]


def test_epoch_with_ragged_last_batch_matches_reference_loop(run_dir):
    """train.py:55-74 trains on every batch of the DataLoader, the ragged last one included (12 dates, batch 8 -> 8 + 4).
    The engine steps the full batch through its captured graph and the ragged one eagerly; the epoch must leave the same
    parameters and the same mean loss as the reference loop (module API + torch.optim.AdamW) on the same batches."""
    import copy
    from raincast_gnn_b200 import train as rc_train
    from raincast_gnn_b200.engine import TrainEngine
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import DataLoader
    from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
    cfg = json.load(open(os.path.join(run_dir, "params.json")))
    dev = torch.device("cuda:0")
    ds = SyntheticEUPPBench(n_dates=12, seed=3)
    loader = DataLoader(ds, batch_size=8, shuffle=False)
    sizes = [b.x.shape[0] for b in loader]
    assert sizes == [8 * 122, 4 * 122]
    torch.manual_seed(1)
    model_a = GNN(35, cfg["gnn_hidden"], cfg["gnn_hidden"], cfg["gnn_layers"], torch.optim.AdamW, {"lr": 1e-3}, cfg["loss"],
                  cfg["grad_u"], cfg["u"], cfg["xi"]).to(dev).train()
    model_b = copy.deepcopy(model_a)
    # reference loop
    opt = torch.optim.AdamW(model_a.parameters(), lr=1e-3)
    ref_losses = []
    for b in loader:
        b = b.to(dev)
        loss = model_a.loss_fn.crps(model_a(b), b.y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_losses.append(float(loss))
    # engine epoch
    first = next(iter(loader))
    eng = TrainEngine(model_b, first.station_graph, first.x.shape[0], 11, 35, lr=1e-3).capture()
    mean_loss = rc_train.run_epoch_engine(eng, loader)
    assert int(eng.step_count) == 2
    assert abs(mean_loss - sum(ref_losses) / 2) < 1e-5 * abs(sum(ref_losses) / 2)
    for (k, pa), (_, pb) in zip(model_a.named_parameters(), model_b.named_parameters()):
        # Adam's first steps move every element by ~lr * m / sqrt(v): a skipped or doubled step would shift almost
        # every element by ~1e-3 = lr.  The normalisation amplifies the RELATIVE error of an element's gradient, and fp32
        # kernels are only max-norm accurate, so single elements with a small gradient may differ by a few percent of
        # lr (all of them where the true gradient is 0: the bias in front of BatchNorm); the mean over a tensor may not.
        d = (pa - pb).abs().flatten()
        assert d.max().item() < 0.5e-3, k
        if not k.endswith(".nn.0.bias"):
            assert d.mean().item() < 0.02 * 1e-3, k
