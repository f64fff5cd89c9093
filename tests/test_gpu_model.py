"""GPU parity of the whole hot path through the reference-facing module API (GNN, loss_fn.crps,
backward, state_dict), against the fixtures written by the reference's own modules and against the
CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from conftest import grad_scale, rel_err
from oracle.make_golden import MODEL_CASES, model_case_inputs, summarize
from raincast_gnn_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def build_case(name, dev):
    from raincast_gnn_b200.graph import radius_graph
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import Batch
    from raincast_gnn_b200.utils.data import make_graphs
    c = model_case_inputs(name)
    ei, ea = radius_graph(c["dist"], c["max_dist"])
    batch = Batch.from_data_list(make_graphs(c["x"], c["ensemble"], c["y"], ei, ea, c["n"]))
    model = GNN(in_channels=c["f"], hidden_channels_gnn=c["h"], out_channels_gnn=c["h"], num_layers_gnn=c["layers"],
                optimizer_class=torch.optim.AdamW, optimizer_params={"lr": 1e-4}, loss=c["loss"], grad_u=c["grad_u"],
                u=1.71, xi=0.5)
    sd = syn.seeded_state_dict(model.state_dict(), seed=1234)
    model.load_state_dict(sd)
    return c, batch.to(dev), model.to(dev), sd


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_train_step_matches_reference_fixture(dev, golden_model, name):
    c, batch, model, sd = build_case(name, dev)
    assert list(model.state_dict().keys()) == list(golden_model[f"{name}.keys"])
    model.train()
    preds = model(batch)
    loss = model.loss_fn.crps(preds, batch.y)
    loss.backward()
    assert rel_err(preds.detach().cpu().numpy(), golden_model[f"{name}.train.preds"]) < TOL
    want = float(golden_model[f"{name}.train.loss"])
    assert abs(loss.item() - want) < TOL * abs(want)
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    for k, gr in grads.items():
        assert gr.shape == dict(model.named_parameters())[k].shape
        if f"{name}.grad.{k}" in golden_model:
            ref = golden_model[f"{name}.grad.{k}"]
            scale = grad_scale(k, np.abs(ref).max(), lambda kk: np.abs(golden_model[f"{name}.grad.{kk}"]).max())
            assert np.abs(gr.numpy() - ref).max() / scale < TOL, k
        else:
            ref = golden_model[f"{name}.gradsum.{k}"]
            scale = grad_scale(k, ref[2], lambda kk: golden_model[f"{name}.gradsum.{kk}"][2])
            got = summarize(gr)
            assert abs(got[2] - ref[2]) <= TOL * scale, k
            assert abs(got[3] - ref[3]) <= TOL * scale * np.sqrt(gr.numel()) * 4, k
            assert np.abs(gr.reshape(-1)[:32].numpy() - golden_model[f"{name}.gradhead.{k}"]).max() <= TOL * scale, k
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            assert rel_err(v.cpu().numpy(), golden_model[f"{name}.buf.{k}"]) < TOL, k
    model.eval()
    with torch.no_grad():
        assert rel_err(model(batch).cpu().numpy(), golden_model[f"{name}.eval.preds"]) < TOL


@pytest.mark.parametrize("name", ["tiny_mixed_u", "ref_mixed_u"])
def test_adamw_trajectory_matches_reference_fixture(dev, golden_model, name):
    """train.py:64-69 verbatim (torch.optim.AdamW on the module's parameters) for three steps."""
    c, batch, model, sd = build_case(name, dev)
    model.train()
    opt = model.optimizer_class(model.parameters(), **model.optimizer_params)
    traj = []
    for _ in range(3):
        preds = model(batch)
        loss = model.loss_fn.crps(preds, batch.y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        traj.append(loss.item())
    assert rel_err(np.array(traj), golden_model[f"{name}.adamw.losses"]) < TOL
    assert rel_err(model.aggr.weight.detach().cpu().numpy(), golden_model[f"{name}.adamw.aggr_weight"]) < TOL
    assert rel_err(model.conv.convolutions[0].eps.detach().cpu().numpy(), golden_model[f"{name}.adamw.eps0"]) < 1e-4


def test_reference_shape_batch8_vs_oracle(dev):
    """BASELINE.json config 2 shape (B=8 x 122 stations x 11 members, H=128, L=4, mixed_u) against the oracle
    with a shared state_dict: activations, CRPS, every parameter gradient."""
    from oracle import graph as og, model as om, pyg as opyg
    from raincast_gnn_b200.models import GNN
    from raincast_gnn_b200.pyg_compat import DataLoader
    from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
    torch.set_num_threads(4)
    ds = SyntheticEUPPBench(n_dates=8)
    batch = next(iter(DataLoader(ds, batch_size=8)))
    kw = dict(in_channels=35, hidden_channels_gnn=128, out_channels_gnn=128, num_layers_gnn=4,
              optimizer_class=torch.optim.AdamW, optimizer_params={"lr": 1e-4}, loss="MixedLoss", grad_u="True", u=1.71, xi=0.5)
    ref = om.GNN(**kw)
    sd = syn.seeded_state_dict(ref.state_dict(), seed=99)
    ref.load_state_dict(sd)
    ours = GNN(**kw)
    ours.load_state_dict(sd)
    ours.to(dev)
    ref.train()
    ours.train()
    # the reference runs one no-grad forward on a single un-batched Data before training (train.py:182-183)
    single = ds[0]
    with torch.no_grad():
        ref(opyg.Data(x=single.x, ensemble=single.ensemble, edge_index=single.edge_index, edge_attr=single.edge_attr))
        ours(single.to(dev))
    ob = opyg.Data(x=batch.x, ensemble=batch.ensemble, edge_index=batch.edge_index, edge_attr=batch.edge_attr, y=batch.y)
    p_ref = ref(ob)
    l_ref = ref.loss_fn.crps(p_ref, ob.y)
    l_ref.backward()
    b = batch.to(dev)
    p = ours(b)
    l = ours.loss_fn.crps(p, b.y)
    l.backward()
    assert rel_err(p.detach().cpu().numpy(), p_ref.detach().numpy()) < TOL
    assert abs(l.item() - l_ref.item()) < TOL * abs(l_ref.item())
    ref_grads = {k: v.grad for k, v in ref.named_parameters()}
    for k, v in ours.named_parameters():
        want = ref_grads[k].numpy()
        scale = grad_scale(k, np.abs(want).max(), lambda kk: ref_grads[kk].abs().max().item())
        assert np.abs(v.grad.cpu().numpy() - want).max() / scale < TOL, k
    for k, v in ours.state_dict().items():
        assert rel_err(v.cpu().numpy(), ref.state_dict()[k].numpy()) < TOL, k


def test_checkpoint_roundtrip_and_cpu_rejection(dev, tmp_path):
    """Bare state_dict .ckpt (train.py:207 / eval.py:196-197) loads strictly into the oracle-shaped model and back;
    CPU tensors are rejected loudly (no CPU path)."""
    from oracle import model as om
    from raincast_gnn_b200 import _lib
    c, batch, model, sd = build_case("tiny_mixed_u", dev)
    path = tmp_path / "run_0-best.ckpt"
    torch.save(model.state_dict(), path)
    ck = torch.load(path, map_location="cpu")
    ref = om.GNN(in_channels=c["f"], hidden_channels_gnn=c["h"], out_channels_gnn=c["h"], num_layers_gnn=c["layers"],
                 loss=c["loss"], grad_u=c["grad_u"], u=1.71, xi=0.5)
    ref.load_state_dict(ck, strict=True)
    model.load_state_dict(torch.load(path, map_location=dev), strict=True)
    with pytest.raises(_lib.RcError):
        model.cpu()(batch.to("cpu"))
