"""CUDA-backed mirror of the reference loss classes (models/loss.py): same names, constructor
arguments and `crps(prediction, y)` call (train.py:65,87; eval.py:213), value and gradient from the
single-pass kernel rc_crps_fwd_bwd.  The result is a float64 scalar like the reference's (its
`c = np.log(0.01)` promotes the mean, models/loss.py:72,230-231); NaN targets are skipped
(models/loss.py:216).  CPU tensors are rejected: there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..functional import CrpsFn


class _CrpsBase(torch.nn.Module):
    kind = None

    def _call(self, prediction, y, u=0.0, xi=0.5, t=5.0):
        width = self.kind + 2
        if prediction.dim() != 2 or prediction.shape[1] != width:
            raise ValueError(f"{type(self).__name__}.crps expects prediction of shape [M, {width}], got {tuple(prediction.shape)}")
        home = prediction.device
        if not prediction.is_cuda:
            # eval.py:213 scores checkpoint-averaged predictions that predict_model moved to the host
            # (eval.py:68); they are scored by the same kernel after a copy to the GPU — never on the CPU.
            if not torch.cuda.is_available():
                raise _lib.RcError("crps needs a CUDA device; there is no CPU path")
            prediction = prediction.cuda()
        if y.device != prediction.device:
            y = y.to(prediction.device)
        out = CrpsFn.apply(prediction, y, self.kind, float(u), float(xi), float(t))
        return out if home == prediction.device else out.to(home)


class NormalCRPS(_CrpsBase):
    """models/loss.py:335-369."""
    kind = _lib.RC_LOSS_NORMAL

    def crps(self, prediction: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self._call(prediction, y)


class MixedNormalCRPS(_CrpsBase):
    """models/loss.py:6-68 (censored normal with a point mass p at c = log 0.01)."""
    kind = _lib.RC_LOSS_MIXED_NORMAL

    def __init__(self, reduce: bool = True, c: float = np.log(0.01)):
        super().__init__()
        if not reduce:
            raise NotImplementedError("reduce=False is not used by the reference's training / eval path")
        if abs(float(c) - float(np.log(0.01))) > 1e-12:
            raise NotImplementedError("the censoring point is fixed at log(0.01) (utils/data.py:204)")
        self.reduce, self.c = reduce, c

    def crps(self, prediction: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self._call(prediction, y)


class MixedLoss(_CrpsBase):
    """models/loss.py:71-272: censored-normal body + point mass at c + GPD tail above u
    (u learned: sigmoid blend with t = 5, :266; u fixed: hard switch, :268)."""

    def __init__(self, grad_u: bool, xi: float, u=None, reduce: bool = True, t: float = 5, c=np.log(0.01)):
        super().__init__()
        if not reduce:
            raise NotImplementedError("reduce=False is not used by the reference's training / eval path")
        if abs(float(c) - float(np.log(0.01))) > 1e-12:
            raise NotImplementedError("the censoring point is fixed at log(0.01) (utils/data.py:204)")
        if float(xi) in (0.0, 1.0, 2.0):
            raise ValueError("xi must not be 0, 1 or 2 (models/loss.py:90,121-124 divide by xi, 1-xi, 2-xi)")
        self.reduce, self.c, self.grad_u, self.u, self.xi, self.t = reduce, c, grad_u, u, xi, t
        if not grad_u and u is None:
            raise ValueError("a fixed threshold u is required when grad_u is False")

    @property
    def kind(self):
        return _lib.RC_LOSS_MIXED_U if self.grad_u else _lib.RC_LOSS_MIXED

    def crps(self, prediction: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self._call(prediction, y, u=0.0 if self.grad_u else self.u, xi=self.xi, t=self.t)
