"""torch.autograd.Function wrappers over kernels.py, so that the reference's training loop
(`preds = model(batch); loss = model.loss_fn.crps(preds, batch.y); loss.backward()`, train.py:64-68)
runs unchanged on the CUDA kernels.  Each Function is one block of the network; backward calls the
hand-written backward kernels; the only PyTorch arithmetic is the scalar chain-rule factor of CrpsFn
(grad_output of the loss, 1.0 in train.py).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib
from . import kernels as K


def _prep(t):
    _lib.require_cuda(t)
    return _lib.f32c(t.detach())


class DeepSetsFn(Function):
    """models/gnn.py:64-68."""
    NAMES = ("phi0_w", "phi0_b", "phi2_w", "phi2_b", "rho0_w", "rho0_b", "rho2_w", "rho2_b")

    @staticmethod
    def forward(ctx, ens, bf16, *params):
        P = {n: _prep(p) for n, p in zip(DeepSetsFn.NAMES, params)}
        emb, saved = K.deepsets_fwd(P, _prep(ens), bf16=bf16)
        ctx.P, ctx.saved = P, saved
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        P = ctx.P
        G = {n: torch.empty_like(P[n]) for n in DeepSetsFn.NAMES}
        K.deepsets_bwd(P, ctx.saved, _lib.f32c(d_emb), G)
        return (None, None) + tuple(G[n] for n in DeepSetsFn.NAMES)


class DimRedFn(Function):
    """models/gnn.py:134-135 (cat + Linear as two reduction segments)."""

    @staticmethod
    def forward(ctx, x, emb, w, b):
        P = {"dimred_w": _prep(w), "dimred_b": _prep(b)}
        y, saved = K.dimred_fwd(P, _prep(x), _prep(emb))
        ctx.P, ctx.saved = P, saved
        return y

    @staticmethod
    def backward(ctx, dy):
        P = ctx.P
        G = {n: torch.empty_like(P[n]) for n in P}
        d_emb = K.dimred_bwd(P, ctx.saved, _lib.f32c(dy), G)
        return None, d_emb, G["dimred_w"], G["dimred_b"]


class GineLayerFn(Function):
    """One ResGnn layer: GINEConv(Linear-BN-ReLU-Linear) + ReLU (+ residual), models/gnn.py:21-29,39-44."""
    NAMES = ("eps", "lin_w", "lin_b", "nn0_w", "nn0_b", "bn_w", "bn_b", "nn3_w", "nn3_b")

    @staticmethod
    def forward(ctx, x, graph, first, training, bn_rm, bn_rv, bn_nbt, *params):
        P = {n: _prep(p) for n, p in zip(GineLayerFn.NAMES, params)}
        P.update(bn_rm=bn_rm, bn_rv=bn_rv, bn_nbt=bn_nbt)        # buffers are updated in place by the kernel
        y, saved = K.gine_layer_fwd(P, _prep(x), graph, first=first, training=training)
        ctx.P, ctx.saved, ctx.graph, ctx.first, ctx.training = P, saved, graph, first, training
        return y

    @staticmethod
    def backward(ctx, dy):
        P = ctx.P
        G = {n: torch.empty_like(P[n]) for n in GineLayerFn.NAMES}
        dx = K.gine_layer_bwd(P, ctx.saved, ctx.graph, _lib.f32c(dy), G, first=ctx.first, training=ctx.training)
        return (dx, None, None, None, None, None, None) + tuple(G[n] for n in GineLayerFn.NAMES)


class HeadFn(Function):
    """models/gnn.py:123,139 (`aggr` Linear H -> C)."""

    @staticmethod
    def forward(ctx, x, w, b):
        P = {"aggr_w": _prep(w), "aggr_b": _prep(b)}
        y, saved = K.head_fwd(P, _prep(x))
        ctx.P, ctx.saved = P, saved
        return y

    @staticmethod
    def backward(ctx, dy):
        P = ctx.P
        G = {n: torch.empty_like(P[n]) for n in P}
        dx = K.head_bwd(P, ctx.saved, _lib.f32c(dy), G)
        return dx, G["aggr_w"], G["aggr_b"]


class PostProcessFn(Function):
    """models/model_utils.py:89-113."""

    @staticmethod
    def forward(ctx, raw, kind):
        r = _prep(raw)
        ctx.raw, ctx.kind = r, kind
        return K.postprocess_fwd(r, kind)

    @staticmethod
    def backward(ctx, d_post):
        return K.postprocess_bwd(ctx.raw, _lib.f32c(d_post), ctx.kind), None


class CrpsFn(Function):
    """models/loss.py:203-272 / :12-68 / :346-369: value and gradient from one kernel pass."""

    @staticmethod
    def forward(ctx, pred, y, kind, u, xi, t):
        p = _prep(pred)
        yy = _prep(y).reshape(-1)
        if yy.numel() != p.shape[0]:
            raise ValueError(f"y has {yy.numel()} entries for {p.shape[0]} predictions")
        need = ctx.needs_input_grad[0]
        loss, d_pred, _ = K.crps_fwd_bwd(p, yy, kind, raw_input=False, u=u, xi=xi, t=t, need_grad=need)
        ctx.d_pred = d_pred
        return loss.reshape(())                      # float64 scalar, like the reference's promoted result

    @staticmethod
    def backward(ctx, g):
        return ctx.d_pred * g.to(torch.float32), None, None, None, None, None
