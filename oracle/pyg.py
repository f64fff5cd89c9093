"""Oracle (test infrastructure): restatement of the torch-geometric pieces the
reference calls but does not vendor.

`torch-geometric` is an unpinned pip dependency of the reference
(environment.yml:28-31; README.md:73 says "2.3.1+") and is neither present under
/root/reference nor installable in this image, so its published algorithm is
restated here from the PyG documentation / public source:

  GINEConv (models/gnn.py:5,27-29 call site, train_eps=True, edge_dim=1):
      lin  = Linear(edge_dim, nn[0].in_features)
      m_e  = relu(x[edge_index[0, e]] + lin(edge_attr[e]))
      agg_i = sum_{e: edge_index[1, e] = i} m_e        (aggr='add')
      out  = nn(agg + (1 + eps) * x),   eps: Parameter of shape [1], init 0
  Batch.from_data_list (train.py:155-156, eval.py:141 call sites):
      attributes whose name contains "index" are concatenated on dim -1 and
      shifted by the cumulative node count, every other tensor on dim 0;
      `batch` (graph id per node) and `ptr` are added.

Parity for this file is UNPINNED by the reference (it has no tests).
"""
from __future__ import annotations

import torch
from torch import nn


class GINEConv(nn.Module):
    def __init__(self, nn: nn.Module, eps: float = 0.0, train_eps: bool = False, edge_dim=None):
        super().__init__()
        self.nn = nn
        self.initial_eps = float(eps)
        if train_eps:
            self.eps = torch.nn.Parameter(torch.empty(1))
        else:
            self.register_buffer("eps", torch.empty(1))
        first = nn[0] if isinstance(nn, torch.nn.Sequential) else nn
        in_channels = first.in_features
        self.lin = torch.nn.Linear(edge_dim, in_channels) if edge_dim is not None else None
        with torch.no_grad():
            self.eps.fill_(self.initial_eps)

    def forward(self, x, edge_index, edge_attr):
        src, dst = edge_index[0], edge_index[1]
        msg = (x.index_select(0, src) + self.lin(edge_attr)).relu()
        agg = torch.zeros_like(x).index_add_(0, dst, msg)
        return self.nn(agg + (1 + self.eps) * x)


class Data:
    """Attribute bag with the handful of PyG `Data` behaviours the reference uses."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device):
        out = type(self)()
        for k in self.keys():
            v = getattr(self, k)
            setattr(out, k, v.to(device) if torch.is_tensor(v) else v)
        return out

    def clone(self):
        out = type(self)()
        for k in self.keys():
            v = getattr(self, k)
            setattr(out, k, v.clone() if torch.is_tensor(v) else v)
        return out

    @property
    def num_nodes(self):
        return self.x.shape[0]


class Batch(Data):
    @classmethod
    def from_data_list(cls, items):
        out = cls()
        keys = items[0].keys()
        counts = [d.num_nodes for d in items]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        for k in keys:
            vals = [getattr(d, k) for d in items]
            if not torch.is_tensor(vals[0]):
                setattr(out, k, vals)
            elif "index" in k:
                setattr(out, k, torch.cat([v + o for v, o in zip(vals, offs)], dim=-1))
            else:
                setattr(out, k, torch.cat(vals, dim=0))
        out.batch = torch.repeat_interleave(torch.arange(len(items)), torch.tensor(counts))
        out.ptr = torch.tensor(offs, dtype=torch.long)
        return out


class DataLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n).tolist() if self.shuffle else list(range(n))
        for i in range(0, n, self.batch_size):
            yield Batch.from_data_list([self.dataset[j] for j in order[i:i + self.batch_size]])
