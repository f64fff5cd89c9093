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
    pin_outputs = False        # collate node tensors straight into pinned host memory (set by train.run_epoch_engine while it runs)

    @classmethod
    def from_data_list(cls, items):
        """torch_geometric.data.Batch.from_data_list for the attributes the reference uses (utils/data.py:287-340): node
        tensors concatenated, `edge_index` offset per graph, `batch` / `ptr`.  With the reference's static station graph
        (every item shares ONE edge_index / edge_attr tensor, utils/data.py:300) the collated edge tensors, `batch`, `ptr`
        and the CSR layout are the same for every batch of that size: they are built once and shared (read-only) by the
        batches - per step only x / ensemble / y are concatenated, into pinned memory when `pin_outputs` is set."""
        out = cls()
        counts = [d.num_nodes for d in items]
        e0, a0 = getattr(items[0], "edge_index", None), getattr(items[0], "edge_attr", None)
        shared = (e0 is not None and a0 is not None and len(set(counts)) == 1 and
                  all(getattr(d, "edge_index", None) is e0 and getattr(d, "edge_attr", None) is a0 for d in items))
        key = (e0.data_ptr(), a0.data_ptr(), e0._version, len(items), counts[0]) if shared else None
        hit = cls._GRAPH_CACHE.get(key) if key is not None else None
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        pin = cls.pin_outputs and torch.cuda.is_available()
        for k in items[0].keys():
            vals = [getattr(d, k) for d in items]
            if not torch.is_tensor(vals[0]):
                setattr(out, k, vals)
            elif hit is not None and k in ("edge_index", "edge_attr"):
                setattr(out, k, hit[3][k])
            elif "index" in k:
                setattr(out, k, torch.cat([v + o for v, o in zip(vals, offs)], dim=-1))
            elif pin and vals[0].device.type == "cpu":
                shape = (sum(v.shape[0] for v in vals),) + tuple(vals[0].shape[1:])
                buf = torch.empty(shape, dtype=vals[0].dtype, pin_memory=True)
                setattr(out, k, torch.cat(vals, dim=0, out=buf))
            else:
                setattr(out, k, torch.cat(vals, dim=0))
        if hit is not None:
            out.batch, out.ptr, out.station_graph = hit[3]["batch"], hit[3]["ptr"], hit[0]
            return out
        out.batch = torch.repeat_interleave(torch.arange(len(items)), torch.tensor(counts))
        out.ptr = torch.tensor(offs, dtype=torch.long)
        g = build_station_graph(out.edge_index, out.edge_attr, int(out.x.shape[0]))
        if key is not None:
            if len(cls._GRAPH_CACHE) >= 64:
                cls._GRAPH_CACHE.pop(next(iter(cls._GRAPH_CACHE)))
            cls._GRAPH_CACHE[key] = (g, e0, a0, {"edge_index": out.edge_index, "edge_attr": out.edge_attr, "batch": out.batch, "ptr": out.ptr})
        out.station_graph = g
        return out


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
