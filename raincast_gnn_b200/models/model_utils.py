"""CUDA-backed mirror of models/model_utils.py: the output link functions."""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib
from ..functional import PostProcessFn

EPS = 1e-6

_KIND = {"NormalCRPS": _lib.RC_LOSS_NORMAL, "MixedNormalCRPS": _lib.RC_LOSS_MIXED_NORMAL}


def loss_kind(loss: str, grad_u) -> int:
    """Kernel id of a params.json (loss, grad_u) pair; grad_u is the STRING "True"/"False" there
    (models/model_utils.py:99, models/gnn.py:98)."""
    if loss in _KIND:
        return _KIND[loss]
    if loss == "MixedLoss":
        return _lib.RC_LOSS_MIXED_U if grad_u == "True" else _lib.RC_LOSS_MIXED
    raise ValueError(f"unknown loss {loss!r}")


class PostProcess(nn.Module):
    """models/model_utils.py:70-113: sigma = softplus + 1e-6, p = sigmoid, u = 2.12 * sigmoid."""

    def __init__(self, loss, grad_u):
        super().__init__()
        self.loss, self.grad_u = loss, grad_u
        self.kind = loss_kind(loss, grad_u)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        width = self.kind + 2
        if x.shape[-1] != width:
            raise ValueError(f"PostProcess({self.loss}) expects {width} columns, got {x.shape[-1]}")
        lead = x.shape[:-1]
        return PostProcessFn.apply(x.reshape(-1, width), self.kind).reshape(*lead, width)


class MakePositive(nn.Module):
    """models/model_utils.py:42-68 (the NormalCRPS branch of PostProcess)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return PostProcessFn.apply(x.reshape(-1, 2), _lib.RC_LOSS_NORMAL).reshape(x.shape)
