"""Data / Batch / DataLoader with the PyG semantics the hot path relies on [PyG-upstream]:
tensor attributes whose name contains "index" are concatenated on the last dim and shifted by the
cumulative node count, everything else on dim 0; `batch` and `ptr` are added; non-tensor attributes
become lists.  In addition every Batch carries `station_graph` (raincast_gnn_b200.graph.StationGraph).
"""
from __future__ import annotations

import torch

from ..graph import StationGraph, build_station_graph


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_") and k != "station_graph"]

    @property
    def num_nodes(self):
        return int(self.x.shape[0])

    def _map(self, fn):
        out = type(self)()
        for k in self.keys():
            v = getattr(self, k)
            setattr(out, k, fn(v) if torch.is_tensor(v) else v)
        g = getattr(self, "station_graph", None)
        if g is not None:
            out.station_graph = g
        return out

    def to(self, device, non_blocking: bool = False):
        out = self._map(lambda t: t.to(device, non_blocking=non_blocking))
        g = getattr(self, "station_graph", None)
        if g is not None:
            out.station_graph = _graph_on(g, torch.device(device))
        return out

    def clone(self):
        return self._map(lambda t: t.clone())

    def __repr__(self):
        parts = [f"{k}={list(getattr(self, k).shape) if torch.is_tensor(getattr(self, k)) else '...'}" for k in self.keys()]
        return f"{type(self).__name__}({', '.join(parts)})"


_DEVICE_GRAPHS: dict = {}


def _graph_on(g: StationGraph, device) -> StationGraph:
    """Device copy of a host StationGraph, made once (the graph is static across batches)."""
    if g.device == device:
        return g
    key = (id(g), str(device))
    hit = _DEVICE_GRAPHS.get(key)
    if hit is None:
        hit = (g.to(device), g)
        _DEVICE_GRAPHS[key] = hit
    return hit[0]


class Batch(Data):
    _GRAPH_CACHE: dict = {}

    @classmethod
    def from_data_list(cls, items):
        out = cls()
        counts = [d.num_nodes for d in items]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        for k in items[0].keys():
            vals = [getattr(d, k) for d in items]
            if not torch.is_tensor(vals[0]):
                setattr(out, k, vals)
            elif "index" in k:
                setattr(out, k, torch.cat([v + o for v, o in zip(vals, offs)], dim=-1))
            else:
                setattr(out, k, torch.cat(vals, dim=0))
        out.batch = torch.repeat_interleave(torch.arange(len(items)), torch.tensor(counts))
        out.ptr = torch.tensor(offs, dtype=torch.long)
        out.station_graph = cls._station_graph(items, out, counts)
        return out

    @classmethod
    def _station_graph(cls, items, out, counts):
        # static graph (utils/data.py:300): every item shares the same edge tensors -> one layout per batch size
        e0, a0 = items[0].edge_index, items[0].edge_attr
        shared = all(d.edge_index is e0 and d.edge_attr is a0 for d in items) and len(set(counts)) == 1
        key = (e0.data_ptr(), a0.data_ptr(), e0._version, len(items), counts[0]) if shared else None
        if key is not None and key in cls._GRAPH_CACHE:
            return cls._GRAPH_CACHE[key][0]
        g = build_station_graph(out.edge_index, out.edge_attr, int(out.x.shape[0]))
        if key is not None:
            if len(cls._GRAPH_CACHE) >= 64:
                cls._GRAPH_CACHE.pop(next(iter(cls._GRAPH_CACHE)))
            cls._GRAPH_CACHE[key] = (g, e0, a0)
        return g


class DataLoader:
    """torch_geometric.loader.DataLoader(dataset, batch_size, shuffle) as used at train.py:155-156, eval.py:141
    (num_workers=0: collation on the calling thread; shuffling draws from torch's global RNG)."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, **_unused):
        self.dataset, self.batch_size, self.shuffle = dataset, int(batch_size), bool(shuffle)

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n).tolist() if self.shuffle else list(range(n))
        for i in range(0, n, self.batch_size):
            yield Batch.from_data_list([self.dataset[j] for j in order[i:i + self.batch_size]])
