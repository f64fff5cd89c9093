"""Generate tests/golden/*.npz from the REFERENCE's own modules (build container only).

Run:  python -m oracle.make_golden            (needs /root/reference; never run on the GPU box)

What is imported verbatim from /root/reference:
  models/loss.py, models/model_utils.py, models/gnn.py  (GNN, ResGnn, DeepSetEncoder,
  PostProcess, MixedLoss, MixedNormalCRPS, NormalCRPS) and utils/data.py's
  build_edge_index_and_attr.
What is NOT the reference: `torch_geometric` (absent, not installable).  oracle/pyg.py's
restated GINEConv / Data are injected under that module name so the reference files import;
geopy / xarray (unused on this path) are stubbed.
Inputs and weights are regenerated from seeds on the test side
(raincast_gnn_b200/utils/synthetic.py), so the fixtures hold outputs only.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _install_shims():
    from oracle import pyg
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_nn.GINEConv = pyg.GINEConv
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data = pyg.Data
    tg_data.InMemoryDataset = object
    tg_loader = types.ModuleType("torch_geometric.loader")
    tg_loader.DataLoader = pyg.DataLoader
    tg.nn, tg.data, tg.loader = tg_nn, tg_data, tg_loader
    geopy = types.ModuleType("geopy")
    geopy.distance = types.ModuleType("geopy.distance")
    xarray = types.ModuleType("xarray")
    xarray.Dataset = object
    for name, mod in {"torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.data": tg_data,
                      "torch_geometric.loader": tg_loader, "geopy": geopy, "geopy.distance": geopy.distance,
                      "xarray": xarray}.items():
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = mod
    if REF not in sys.path:
        sys.path.insert(0, REF)


# ----------------------------------------------------------------------------- shared case definitions
# (imported by the tests so that both sides build identical inputs)

def crps_case_inputs(seed: int, n: int, width: int):
    """Raw head outputs [n,width] and targets y [n] that hit the edge cases of SURVEY.md 8c:
    NaN targets, y == c exactly, y above/below u, tiny and large sigma."""
    from raincast_gnn_b200.utils.synthetic import LOG_001, log_precip_targets
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(n, width, generator=g) * 1.5
    raw[:, 0] = raw[:, 0] * 1.5 - 1.0                       # mu around the log-precip range
    raw[::7, 1] = -12.0                                     # softplus -> sigma ~ 6e-6 + 1e-6
    raw[3::11, 1] = 25.0                                    # beyond the softplus threshold (20)
    if width >= 4:
        raw[5::13, 3] = -9.0
    y = log_precip_targets(n, seed=seed)
    y[1::17] = 4.5                                          # far in the GPD tail
    y[2::19] = float(LOG_001)
    return raw, y


MODEL_CASES = {
    # name: (N stations, B graphs, Em, F, H, L, loss, grad_u, max_dist, box)
    "tiny_mixed_u": (10, 3, 4, 7, 32, 2, "MixedLoss", "True", 250.0, 600.0),
    "tiny_mixed": (10, 3, 4, 7, 32, 2, "MixedLoss", "False", 250.0, 600.0),
    "tiny_normal": (9, 2, 3, 5, 32, 1, "NormalCRPS", "False", 250.0, 600.0),
    "tiny_mixednormal": (9, 2, 3, 5, 32, 3, "MixedNormalCRPS", "False", 250.0, 600.0),
    "ref_mixed_u": (122, 2, 11, 35, 128, 4, "MixedLoss", "True", 100.0, 600.0),
}


def model_case_inputs(name: str):
    from raincast_gnn_b200.utils import synthetic as syn
    n, b, em, f, h, layers, loss, grad_u, max_dist, box = MODEL_CASES[name]
    coords = syn.station_coords(n, box, seed=0)
    dist = syn.distance_matrix(coords)
    x, ens = syn.node_features(n * b, em, f, seed=42)
    y = syn.log_precip_targets(n * b, seed=42)
    return dict(n=n, b=b, em=em, f=f, h=h, layers=layers, loss=loss, grad_u=grad_u, max_dist=max_dist,
                dist=dist, x=x, ensemble=ens, y=y)


def probe_vector(numel: int, seed: int = 7) -> torch.Tensor:
    return torch.randn(numel, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


def summarize(t: torch.Tensor) -> np.ndarray:
    """[sum, abs-sum, max-abs, <t, probe>] in float64 — a compact fingerprint for big tensors."""
    d = t.detach().double().reshape(-1)
    return np.array([d.sum().item(), d.abs().sum().item(), d.abs().max().item(),
                     (d * probe_vector(d.numel())).sum().item()])


# ----------------------------------------------------------------------------- generators

def gen_graph():
    import utils.data as ref_data                           # /root/reference/utils/data.py
    from raincast_gnn_b200.utils import synthetic as syn
    out = {}
    cases = {"ref122_d100": (122, 600.0, 100.0), "ref122_d1": (122, 600.0, 1.0),
             "n7_d300": (7, 600.0, 300.0), "n40_d150": (40, 600.0, 150.0)}
    for name, (n, box, md) in cases.items():
        dist = syn.distance_matrix(syn.station_coords(n, box, seed=0))
        ei, ea = ref_data.build_edge_index_and_attr(dist, md)
        out[f"{name}.edge_index"] = ei.numpy()
        out[f"{name}.edge_attr"] = ea.numpy()
    # an asymmetric "distance" matrix (directed graph): exercises the general transpose path
    rng = np.random.default_rng(3)
    dist = rng.uniform(1.0, 400.0, (23, 23)).astype(np.float32)
    ei, ea = ref_data.build_edge_index_and_attr(dist, 120.0)
    out["asym23.dist"] = dist
    out["asym23.edge_index"] = ei.numpy()
    out["asym23.edge_attr"] = ea.numpy()
    np.savez_compressed(os.path.join(OUT, "graph.npz"), **out)
    print("graph:", {k: v.shape for k, v in out.items()})


def gen_crps():
    from models.loss import MixedLoss, MixedNormalCRPS, NormalCRPS
    from models.model_utils import PostProcess
    out = {}
    cfgs = [("mixed_u", "MixedLoss", "True", 5), ("mixed", "MixedLoss", "False", 4),
            ("mixednormal", "MixedNormalCRPS", "False", 3), ("normal", "NormalCRPS", "False", 2)]
    for tag, loss, grad_u, width in cfgs:
        for seed, n in ((11, 257), (12, 64)):
            raw, y = crps_case_inputs(seed, n, width)
            raw = raw.clone().requires_grad_(True)
            post = PostProcess(loss, grad_u)(raw)
            post.retain_grad()
            if loss == "MixedLoss":
                fn = MixedLoss(grad_u=(grad_u == "True"), xi=0.5, u=None if grad_u == "True" else 1.71)
            elif loss == "MixedNormalCRPS":
                fn = MixedNormalCRPS()
            else:
                fn = NormalCRPS()
            val = fn.crps(post, y)
            val.backward()
            key = f"{tag}.s{seed}"
            out[f"{key}.post"] = post.detach().numpy()
            out[f"{key}.loss"] = np.array(val.item(), dtype=np.float64)
            out[f"{key}.loss_dtype"] = np.array(str(val.dtype))
            out[f"{key}.dpost"] = post.grad.numpy()
            out[f"{key}.draw"] = raw.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "crps.npz"), **out)
    print("crps:", len(out), "arrays")


def gen_model():
    import utils.data as ref_data
    from models.gnn import GNN
    from oracle.pyg import Batch, Data
    from raincast_gnn_b200.utils.synthetic import seeded_state_dict
    out = {}
    for name in MODEL_CASES:
        c = model_case_inputs(name)
        ei, ea = ref_data.build_edge_index_and_attr(c["dist"], c["max_dist"])
        n, b = c["n"], c["b"]
        items = [Data(x=c["x"][i * n:(i + 1) * n], ensemble=c["ensemble"][i * n:(i + 1) * n],
                      edge_index=ei, edge_attr=ea, y=c["y"][i * n:(i + 1) * n]) for i in range(b)]
        batch = Batch.from_data_list(items)
        torch.manual_seed(0)
        model = GNN(in_channels=c["f"], hidden_channels_gnn=c["h"], out_channels_gnn=c["h"],
                    num_layers_gnn=c["layers"], optimizer_class=torch.optim.AdamW,
                    optimizer_params={"lr": 1e-4}, loss=c["loss"], grad_u=c["grad_u"], u=1.71, xi=0.5)
        sd = seeded_state_dict(model.state_dict(), seed=1234)
        model.load_state_dict(sd)
        out[f"{name}.keys"] = np.array(list(model.state_dict().keys()))
        # --- train-mode step
        model.train()
        preds = model(batch)
        loss = model.loss_fn.crps(preds, batch.y)
        loss.backward()
        out[f"{name}.train.preds"] = preds.detach().numpy().copy()
        out[f"{name}.train.loss"] = np.array(loss.item(), dtype=np.float64)
        small = name.startswith("tiny")
        for k, p in model.named_parameters():
            if small:
                out[f"{name}.grad.{k}"] = p.grad.numpy().copy()
            else:
                out[f"{name}.gradsum.{k}"] = summarize(p.grad)
                out[f"{name}.gradhead.{k}"] = p.grad.reshape(-1)[:32].numpy().copy()
        for k, v in model.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out[f"{name}.buf.{k}"] = v.numpy().copy()
        # --- eval-mode forward with the updated running statistics
        model.eval()
        with torch.no_grad():
            out[f"{name}.eval.preds"] = model(batch).numpy()
        # --- three AdamW steps (train.py:64-69), loss trajectory + final parameter fingerprint
        model.train()
        model.load_state_dict(sd)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        traj = []
        for _ in range(3):
            loss = model.loss_fn.crps(model(batch), batch.y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            traj.append(loss.item())
        out[f"{name}.adamw.losses"] = np.array(traj, dtype=np.float64)
        out[f"{name}.adamw.aggr_weight"] = model.aggr.weight.detach().numpy().copy()
        out[f"{name}.adamw.eps0"] = model.conv.convolutions[0].eps.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "model.npz"), **out)
    print("model:", len(out), "arrays")


def main():
    torch.set_num_threads(1)            # bit-stable sums for the fixtures
    os.makedirs(OUT, exist_ok=True)
    _install_shims()
    gen_graph()
    gen_crps()
    gen_model()


if __name__ == "__main__":
    main()
