"""Oracle (test infrastructure): the network of models/gnn.py with every ReLU decision supplied from outside.

Gradients of a ReLU network are discontinuous in the pre-activations: a unit whose pre-activation sits within
fp32 rounding of zero is on in one correct evaluation and off in another, and whichever way it falls moves whole
gradient tensors by 1e-4 .. 3e-3 of their scale at the reference shape.  Instead of allowing for such flips, the
GPU parity tests read the decisions the CUDA backward took (include/rc_b200.h, "Test instrumentation") and evaluate
THIS restatement with those decisions forced: relu(z) is replaced by z * mask.  With the masks fixed the network is
a smooth function of its inputs and parameters, autograd differentiates it as such, and every gradient tensor can be
held to the north_star tolerance (1e-5) without an allowance.

With `masks=None` every mask is computed from the restatement's own pre-activations (z > 0) and the function is the
plain network: tests/test_oracle_pinned.py pins it to oracle.model.GNN (which is pinned to the reference fixtures).

Restated: DeepSetEncoder (models/gnn.py:48-68), GNN.forward (:129-141), ResGnn.forward (:35-45), the node MLP
Linear-BatchNorm1d(train)-ReLU-Linear (:21-26), PyG GINEConv (message relu(x_j + lin(e)), 'add' aggregation,
(1 + eps) x_i + aggr; oracle/pyg.py).  The ReLU sites, in forward order:
    "phi"   [M, Em, H]   member MLP                         models/gnn.py:53
    "rho"   [M, H]                                          models/gnn.py:59
    "msg{i}" [E, H]      GINE message of layer i, reference edge order
    "bn{i}"  [M, H]      ReLU behind BatchNorm              models/gnn.py:24
    "out{i}" [M, H]      ReLU on the layer output           models/gnn.py:42,44
"""
from __future__ import annotations

import torch

from . import losses

BN_EPS = 1e-5


def _relu(z, masks, used, key):
    m = masks[key].to(z.device) if masks is not None and key in masks else (z > 0)
    used[key] = m
    return z * m.to(z.dtype)


def forward_raw(params: dict, x, ensemble, edge_index, edge_attr, num_layers: int, masks=None):
    """Raw head output [M, C] in train mode (BatchNorm batch statistics) and the dict of masks used.
    `params`: the reference state_dict keys -> tensors (leaves requiring grad for a gradient evaluation)."""
    used = {}
    P = params
    h = ensemble @ P["deepset.phi.0.weight"].T + P["deepset.phi.0.bias"]
    h = _relu(h, masks, used, "phi")
    h = h @ P["deepset.phi.2.weight"].T + P["deepset.phi.2.bias"]
    pooled = h.sum(dim=1)                                                  # SUM over members, models/gnn.py:67
    r = _relu(pooled @ P["deepset.rho.0.weight"].T + P["deepset.rho.0.bias"], masks, used, "rho")
    emb = r @ P["deepset.rho.2.weight"].T + P["deepset.rho.2.bias"]
    node = torch.cat([x, emb], dim=1) @ P["dim_red.weight"].T + P["dim_red.bias"]
    src, dst = edge_index[0], edge_index[1]
    xx = node
    for i in range(num_layers):
        pre = f"conv.convolutions.{i}."
        e = edge_attr @ P[pre + "lin.weight"].T + P[pre + "lin.bias"]
        msg = _relu(xx.index_select(0, src) + e, masks, used, f"msg{i}")
        agg = torch.zeros_like(xx).index_add_(0, dst, msg)
        hh = agg + (1 + P[pre + "eps"]) * xx
        t = hh @ P[pre + "nn.0.weight"].T + P[pre + "nn.0.bias"]
        mean = t.mean(dim=0)
        var = t.var(dim=0, unbiased=False)
        z = (t - mean) / torch.sqrt(var + BN_EPS) * P[pre + "nn.1.weight"] + P[pre + "nn.1.bias"]
        u = _relu(z, masks, used, f"bn{i}")
        o = u @ P[pre + "nn.3.weight"].T + P[pre + "nn.3.bias"]
        y = _relu(o, masks, used, f"out{i}")
        xx = y if i == 0 else xx + y
    raw = xx @ P["aggr.weight"].T + P["aggr.bias"]
    return raw, used


def loss_and_grads(state_dict: dict, batch, *, num_layers: int, loss: str, grad_u: str, u: float, xi: float, masks=None,
                   dtype=torch.float64, device="cpu"):
    """(preds, loss, {param name: gradient}, masks used): one train-mode forward + CRPS + backward of the restatement in
    `dtype`, with the ReLU decisions of `masks` forced (missing keys / None: the restatement's own)."""
    params = {k: v.detach().to(device=device, dtype=dtype).requires_grad_(True) for k, v in state_dict.items()
              if v.dtype.is_floating_point and "running_" not in k}
    x = batch.x.to(device=device, dtype=dtype)
    ens = batch.ensemble.to(device=device, dtype=dtype)
    ea = batch.edge_attr.to(device=device, dtype=dtype).reshape(-1, 1)
    ei = batch.edge_index.to(device)
    y = batch.y.to(device=device, dtype=dtype)
    raw, used = forward_raw(params, x, ens, ei, ea, num_layers, masks)
    preds = losses.postprocess(raw, loss, grad_u)
    if loss == "MixedLoss":
        learn_u = grad_u == "True"
        val = losses.mixed_loss_crps(preds, y, grad_u=learn_u, xi=xi, u=None if learn_u else u)
    elif loss == "MixedNormalCRPS":
        val = losses.mixed_normal_crps(preds, y)
    else:
        val = losses.normal_crps(preds, y)
    val.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    return preds.detach(), val.detach(), grads, used
