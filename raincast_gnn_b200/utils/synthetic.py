"""Synthetic EUPPBench-shaped data.  Pure numpy/torch, CPU, seeded.

The real dataset (Zenodo zip -> Zarr -> pandas, utils/dataset.py, utils/data.py:19-411)
is not reachable without network; every BASELINE.json config is quoted on
synthetic data "of the reference station/member shape".  Reference shapes:
N=122 stations, Em=11 (reforecast) / 51 (forecast) members, F=35 features.
"""
from __future__ import annotations

import math

import numpy as np
import torch

LOG_001 = float(np.log(0.01))      # log-precip floor, utils/data.py:204


def station_coords(num_nodes: int = 122, box: float = 600.0, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).uniform(0.0, box, (num_nodes, 2))


def distance_matrix(coords: np.ndarray) -> np.ndarray:
    diff = coords[:, None, :] - coords[None, :, :]
    return np.sqrt((diff * diff).sum(-1)).astype(np.float32)


def scaled_graph_radius(num_nodes: int, box: float, mean_degree: float = 29.0) -> float:
    """Radius giving ~`mean_degree` neighbours for uniform points (config 4: 100k nodes, box 1000)."""
    return box * math.sqrt(mean_degree / (math.pi * num_nodes))


def node_features(num_nodes: int, members: int, feats: int, seed: int = 42):
    """x [M,F], ensemble [M,Em,F] ~ N(0,1) (reference data are StandardScaler-ed, utils/data.py:393-399)."""
    g = torch.Generator().manual_seed(seed)
    ens = torch.randn(num_nodes, members, feats, generator=g)
    x = ens[:, 0, :].clone()          # first member's features per station, utils/data.py:318-319
    return x, ens


def log_precip_targets(num_nodes: int, seed: int = 42, p_dry: float = 0.45, p_nan: float = 0.03):
    """y [M]: 45 % exactly log(0.01), else log(Gamma(0.7, 4 mm) + 0.01); 3 % NaN (utils/data.py:204)."""
    rng = np.random.default_rng(seed + 1)
    wet = np.log(rng.gamma(0.7, 4.0, num_nodes) + 0.01)
    y = np.where(rng.uniform(size=num_nodes) < p_dry, LOG_001, wet).astype(np.float32)
    y[rng.uniform(size=num_nodes) < p_nan] = np.nan
    return torch.from_numpy(y)


def seeded_state_dict(template: dict, seed: int = 1234, scale: float = 1.0) -> dict:
    """Deterministic weights for a state_dict with the given keys/shapes (same on every box).

    Linear weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) * scale, biases small, BN affine near
    (1, 0), running_var > 0, eps ~ 0.1.  Used so that fixtures need not store weights.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key in sorted(template.keys()):
        ref = template[key]
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros((), dtype=torch.long)
        elif key.endswith("running_var"):
            out[key] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif key.endswith("running_mean"):
            out[key] = torch.randn(shape, generator=g) * 0.1
        elif key.endswith(".eps"):
            out[key] = torch.randn(shape, generator=g) * 0.1
        elif ".nn.1." in key and key.endswith("weight"):      # BatchNorm gamma
            out[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            out[key] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 2:
            bound = scale / math.sqrt(shape[1])
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            out[key] = 0.1 * torch.randn(shape, generator=g)
    return out
