"""Minimal stand-in for the three torch_geometric names the reference's callers import
(`torch_geometric.loader.DataLoader` train.py:21 / eval.py:20, `torch_geometric.data.Data`
utils/data.py:7).  PyG is not installable in this image; more importantly the batching seam is where
the B200 path differs: `Batch.from_data_list` attaches a dst-sorted CSR `StationGraph` built ONCE per
distinct (static graph, batch size) instead of re-collating `edge_index` every step.

    from raincast_gnn_b200.pyg_compat import install; install()   # makes `import torch_geometric...` resolve here
"""
from __future__ import annotations

import importlib.machinery
import sys
import types

from .data import Batch, Data, DataLoader   # noqa: F401


def install():
    """Register `torch_geometric`, `.data`, `.loader` aliases unless the real package is importable."""
    try:
        import torch_geometric  # noqa: F401
        return False
    except ImportError:
        pass
    from . import data as _data
    tg = types.ModuleType("torch_geometric")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_loader = types.ModuleType("torch_geometric.loader")
    tg_data.Data, tg_data.Batch = _data.Data, _data.Batch
    tg_loader.DataLoader = _data.DataLoader
    tg.data, tg.loader = tg_data, tg_loader
    for name, mod in (("torch_geometric", tg), ("torch_geometric.data", tg_data), ("torch_geometric.loader", tg_loader)):
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = mod
    return True
