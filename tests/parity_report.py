"""Tolerance attribution (not a test): CUDA path vs fp32 oracle vs fp64 oracle on the reference shape.
    python tests/parity_report.py [B] [Em]
Prints, per tensor, max|a-b|/max|b| for (ours, oracle32), (ours, oracle64), (oracle32, oracle64) and the number of
entries of (ours - oracle32) above 1e-5 of the tensor scale."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om, pyg as opyg                       # noqa: E402
from raincast_gnn_b200.models import GNN                          # noqa: E402
from raincast_gnn_b200.pyg_compat import DataLoader               # noqa: E402
from raincast_gnn_b200.utils import synthetic as syn              # noqa: E402
from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench    # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    em = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    dev = torch.device("cuda:0")
    ds = SyntheticEUPPBench(n_dates=B, members=em)
    batch = next(iter(DataLoader(ds, batch_size=B)))
    kw = dict(in_channels=35, hidden_channels_gnn=128, out_channels_gnn=128, num_layers_gnn=4, optimizer_class=None,
              optimizer_params=None, loss="MixedLoss", grad_u="True", u=1.71, xi=0.5)
    ref32 = om.GNN(**kw)
    sd = syn.seeded_state_dict(ref32.state_dict(), seed=99)
    ref32.load_state_dict(sd)
    ref64 = om.GNN(**kw).double()
    ref64.load_state_dict(sd)
    ref64.conv.force_float = False
    ours = GNN(**kw)
    ours.load_state_dict(sd)
    ours.to(dev)
    out = {}
    for tag, model, cast in (("o32", ref32, torch.float32), ("o64", ref64, torch.float64)):
        model.train()
        ob = opyg.Data(x=batch.x.to(cast), ensemble=batch.ensemble.to(cast), edge_index=batch.edge_index,
                       edge_attr=batch.edge_attr.to(cast), y=batch.y.to(cast))
        p = model(ob)
        l = model.loss_fn.crps(p, ob.y)
        l.backward()
        out[tag] = {"preds": p.detach().double(), "loss": l.detach().double().reshape(1),
                    **{k: v.grad.double() for k, v in model.named_parameters()}}
    ours.train()
    b = batch.to(dev)
    p = ours(b)
    l = ours.loss_fn.crps(p, b.y)
    l.backward()
    out["us"] = {"preds": p.detach().cpu().double(), "loss": l.detach().cpu().double().reshape(1),
                 **{k: v.grad.cpu().double() for k, v in ours.named_parameters()}}

    def rel(a, b, scale):
        return float((a - b).abs().max() / scale)

    # float64 gradients at inputs perturbed by fp32-rounding-sized noise: the ReLU-threshold sensitivity
    import copy
    sens = {k: 0.0 for k in out["o64"]}
    for t in range(2):
        g = torch.Generator().manual_seed(1000 + t)
        ref64.zero_grad()
        ob = opyg.Data(x=(batch.x * (1 + 1e-6 * torch.randn(batch.x.shape, generator=g))).double(),
                       ensemble=(batch.ensemble * (1 + 1e-6 * torch.randn(batch.ensemble.shape, generator=g))).double(),
                       edge_index=batch.edge_index, edge_attr=batch.edge_attr.double(), y=batch.y.double())
        p = ref64(ob)
        l = ref64.loss_fn.crps(p, ob.y)
        l.backward()
        pert = {"preds": p.detach(), "loss": l.detach().reshape(1), **{k: v.grad for k, v in ref64.named_parameters()}}
        for k in sens:
            sens[k] = max(sens[k], float((pert[k] - out["o64"][k]).abs().max()))
    print(f"{'tensor':44s} {'us-o32':>9s} {'us-o64':>9s} {'o32-o64':>9s} {'sens64':>9s} {'n>1e-5':>7s} {'numel':>7s}")
    for k in out["us"]:
        scale = max(float(out["o64"][k].abs().max()), 1e-30)
        if k.endswith(".nn.0.bias"):
            scale = float(out["o64"][k[:-4] + "weight"].abs().max())
        d = (out["us"][k] - out["o32"][k]).abs() / scale
        print(f"{k:44s} {rel(out['us'][k], out['o32'][k], scale):9.2e} {rel(out['us'][k], out['o64'][k], scale):9.2e} "
              f"{rel(out['o32'][k], out['o64'][k], scale):9.2e} {sens[k] / scale:9.2e} {int((d > 1e-5).sum()):7d} {d.numel():7d}")


if __name__ == "__main__":
    main()
