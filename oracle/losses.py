"""Oracle (test infrastructure): closed-form CRPS losses and output links, CPU torch.

Restates models/loss.py (MixedLoss :71-272, MixedNormalCRPS :6-68, NormalCRPS
:335-369) and models/model_utils.py (PostProcess :70-113, MakePositive :42-68)
with the common sub-expressions of SURVEY.md Appendix A written once.  Autograd
supplies the gradients, so the conventions to copy (|x|' = 0 at 0, `where`
passes gradient to the selected branch only, softplus threshold 20) are torch's.

dtype quirk kept on purpose (models/loss.py:72,230-231):
the censoring point c = np.log(0.01) is an np.float64, so the reference builds a
float64 one-element tensor from it and every term that touches c is promoted;
the returned loss is float64.  `quirk_f64=False` evaluates in the input dtype.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_001 = np.log(0.01)          # models/loss.py:72
LINK_EPS = 1e-6                 # models/model_utils.py:5
U_SCALE = 2.12                  # models/model_utils.py:104
_SQRT2 = math.sqrt(2.0)
_INV_SQRT_PI = 1.0 / math.sqrt(math.pi)


def _Phi(z):
    return 0.5 * (1 + torch.erf(z / _SQRT2))          # torch Normal(0,1).cdf


def _phi(z):
    # Normal(0,1).log_prob(z).exp(), models/loss.py:148
    return torch.exp(-0.5 * z * z - math.log(math.sqrt(2 * math.pi)))


def postprocess(raw: torch.Tensor, loss: str, grad_u) -> torch.Tensor:
    """models/model_utils.py:89-113.  `grad_u` is the params.json STRING ("True"/"False")."""
    cols = list(torch.split(raw, 1, dim=-1))
    if loss == "NormalCRPS":
        cols[1] = F.softplus(cols[1]) + LINK_EPS
    elif loss == "MixedNormalCRPS":
        cols[1] = F.softplus(cols[1]) + LINK_EPS
        cols[2] = torch.sigmoid(cols[2])
    elif loss == "MixedLoss":
        cols[1] = F.softplus(cols[1]) + LINK_EPS
        cols[2] = torch.sigmoid(cols[2])
        cols[3] = F.softplus(cols[3]) + LINK_EPS
        if grad_u == "True":
            cols[4] = torch.sigmoid(cols[4]) * U_SCALE
    else:
        raise ValueError(loss)
    return torch.cat(cols, dim=-1)


def _censor_point(y, quirk_f64):
    if quirk_f64:
        return torch.tensor([LOG_001], dtype=torch.float64, device=y.device)
    return torch.tensor([float(LOG_001)], dtype=y.dtype, device=y.device)


def normal_crps(pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """models/loss.py:346-369."""
    ok = ~torch.isnan(y)
    mu, sg = pred[:, 0:1][ok.unsqueeze(1)], pred[:, 1:2][ok.unsqueeze(1)]
    z = (y[ok] - mu) / sg
    inv_sqrt_pi = 1 / torch.sqrt(torch.tensor(np.pi))          # float32 constant, :343
    return (sg * (z * (2.0 * _Phi(z) - 1.0) + 2.0 * _phi(z) - inv_sqrt_pi)).mean()


def mixed_normal_crps(pred: torch.Tensor, y: torch.Tensor, quirk_f64: bool = True) -> torch.Tensor:
    """models/loss.py:12-68 (censored normal with a point mass p at c)."""
    ok = ~torch.isnan(y)
    m = ok.unsqueeze(1)
    mu, sg, p = pred[:, 0:1][m], pred[:, 1:2][m], pred[:, 2:3][m]
    yv = y[ok]
    c = _censor_point(yv, quirk_f64)
    zy, zc = (yv - mu) / sg, (c - mu) / sg
    q = 1 - p
    Pc = p + q * _Phi(zc)
    t1 = zy * (2 * (p + q * _Phi(zy)) - 1)
    t2 = -zc * Pc ** 2
    t3 = -2 * q * _phi(zc) * Pc
    t4 = 2 * q * _phi(zy)
    t5 = -(q ** 2) * _INV_SQRT_PI * (1 - _Phi(_SQRT2 * zc))
    return (sg * (t1 + t2 + t3 + t4 + t5)).mean()


def mixed_loss_crps(pred: torch.Tensor, y: torch.Tensor, *, grad_u: bool, xi: float,
                    u: float | None = None, t: float = 5.0, quirk_f64: bool = True,
                    reduce: bool = True) -> torch.Tensor:
    """models/loss.py:203-272 with helpers :81-200."""
    ok = ~torch.isnan(y)
    m = ok.unsqueeze(1)
    mu, sg, p, su = (pred[:, i:i + 1][m] for i in range(4))
    yv = y[ok]
    if grad_u:
        uu = pred[:, 4:5][m]
    else:
        uu = torch.tensor([u], dtype=yv.dtype, device=yv.device)   # :223
    c = _censor_point(yv, quirk_f64)
    q = 1 - p
    zc, zu, zy = (c - mu) / sg, (uu - mu) / sg, (yv - mu) / sg    # :244-246
    Pc = p + q * _Phi(zc)                                          # :141
    Pu = q * (1 - _Phi(zu))                                        # :142
    m_u = p + q * _Phi(zu)                                         # :107
    # t2 + t3 + t5 shared by both normal pieces (:146-160 == :182-198)
    A = (-zc * Pc ** 2 + zu * Pu ** 2
         - 2 * q * _phi(zc) * Pc - 2 * q * _phi(zu) * Pu
         - (q ** 2) * _INV_SQRT_PI * (_Phi(_SQRT2 * zu) - _Phi(_SQRT2 * zc)))
    body = sg * (zy * (2 * (p + q * _Phi(zy)) - 1) + 2 * q * _phi(zy) + A)       # :145,151,162
    upper = sg * (zu + 2 * q * _phi(zu) - 2 * zu * Pu + A)                        # :181,187-189,200

    def tail(v):                                                                   # :111-125
        x = (v - uu) / su
        cdf = torch.where(x <= 0, torch.zeros((), dtype=x.dtype), 1 - (1 + xi * x).pow(-1 / xi))   # :90
        return su * (x.abs() - 2 * (1 - m_u) / (1 - xi) * (1 - (1 - cdf).pow(1 - xi))
                     + (1 - m_u) ** 2 / (2 - xi))

    loss_1 = body + tail(uu)                                       # :250-258
    loss_2 = tail(yv) + upper                                      # :259-263
    if grad_u:
        out = torch.sigmoid((uu - yv) * t) * (loss_1 - loss_2) + loss_2           # :266
    else:
        out = torch.where(yv < uu, loss_1, loss_2)                                 # :268
    return out.mean() if reduce else out


class MixedLoss(torch.nn.Module):
    """Same constructor/`crps` surface as models/loss.py:71-79,203."""

    def __init__(self, grad_u: bool, xi: float, u=None, reduce: bool = True, t: float = 5, c=LOG_001):
        super().__init__()
        self.grad_u, self.xi, self.u, self.reduce, self.t, self.c = grad_u, xi, u, reduce, t, c

    def crps(self, prediction, y):
        return mixed_loss_crps(prediction, y, grad_u=bool(self.grad_u), xi=self.xi, u=self.u,
                               t=self.t, reduce=self.reduce)


class MixedNormalCRPS(torch.nn.Module):
    def crps(self, prediction, y):
        return mixed_normal_crps(prediction, y)


class NormalCRPS(torch.nn.Module):
    def crps(self, prediction, y):
        return normal_crps(prediction, y)
