"""Dataset seam.  `EUPPBench` keeps the reference constructor (utils/dataset.py:29-38) and loads the
processed split files the reference writes (`torch.save((data, slices))`, utils/dataset.py:174-182) when
they exist; it cannot download / process raw Zarr data (no network; xarray/zarr/geopy are not in the
image) and says so.  `SyntheticEUPPBench` generates the SURVEY.md 8d data of the same shapes."""
from __future__ import annotations

import os
from typing import List

import torch

from ..graph import radius_graph
from ..pyg_compat.data import Data
from . import synthetic as syn
from .data import make_graphs


class SyntheticEUPPBench:
    """`n_dates` station graphs of the reference shape sharing one radius graph."""

    def __init__(self, n_dates: int = 64, num_stations: int = 122, members: int = 11, feats: int = 35,
                 max_dist: float = 100.0, box: float = 600.0, seed: int = 42):
        coords = syn.station_coords(num_stations, box, seed=0)
        self.edge_index, self.edge_attr = radius_graph(syn.distance_matrix(coords), max_dist)
        x, ens = syn.node_features(n_dates * num_stations, members, feats, seed=seed)
        y = syn.log_precip_targets(n_dates * num_stations, seed=seed)
        self.graphs: List[Data] = make_graphs(x, ens, y, self.edge_index, self.edge_attr, num_stations)

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, i):
        return self.graphs[i]


class EUPPBench:
    available_splits = ["train_rf", "test_rf", "test_f"]

    def __init__(self, root_raw: str, root_processed: str, leadtime: str = "24h", max_dist: float = 100.0,
                 split: str = "train_rf", transform=None, pre_transform=None):
        if split not in self.available_splits:
            raise ValueError(f"split must be one of {self.available_splits}, got {split}")
        self.leadtime, self.max_dist, self.split = leadtime, max_dist, split
        self.root_raw, self.root_processed = root_raw, root_processed
        path = os.path.join(root_processed, f"EUPPBench_{leadtime}_{split}.pt")
        if not os.path.isfile(path):
            raise FileNotFoundError(
                f"{path} not found. Processing raw EUPPBench Zarr archives (utils/dataset.py:95-182) needs network "
                "access and xarray/zarr/geopy, which this build does not have; use SyntheticEUPPBench or copy the "
                "reference's processed .pt files here.")
        from ..pyg_compat import unpickle                            # the file names torch_geometric classes: read as attribute bags
        data, slices = torch.load(path, weights_only=False, pickle_module=unpickle)
        self.graphs = self._unpack(data, slices)

    @staticmethod
    def _unpack(data, slices) -> List[Data]:
        from ..pyg_compat import unpickle
        attrs = unpickle.attribute_dict(data)
        get = lambda k: attrs[k]                                     # noqa: E731
        n = len(slices["x"]) - 1
        e0, e1 = int(slices["edge_index"][0]), int(slices["edge_index"][1])
        edge_index = get("edge_index")[:, e0:e1].contiguous()       # static graph: share the first copy
        edge_attr = get("edge_attr")[int(slices["edge_attr"][0]):int(slices["edge_attr"][1])].contiguous()
        out = []
        for i in range(n):
            d = Data(**{k: get(k)[int(slices[k][i]):int(slices[k][i + 1])] for k in ("x", "ensemble", "y")},
                     edge_index=edge_index, edge_attr=edge_attr)
            out.append(d)
        return out

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, i):
        return self.graphs[i]


class DeviceSplit:
    """A split held on the GPU: x [D, N, F], ensemble [D, N, Em, F], y [D, N] stacked over the forecast
    dates of a dataset whose graphs share one static station graph.  `TrainEngine.load_dates(split, dates)` then builds
    a batch with one gather kernel and no host work - the reference collates on the CPU main thread and issues seven
    H2D copies per step (train.py:61-62,155-156)."""

    def __init__(self, graphs, device):
        graphs = list(graphs)
        if not graphs:
            raise ValueError("DeviceSplit needs at least one graph")
        first = graphs[0]
        for g in graphs:
            if g.x.shape != first.x.shape or g.ensemble.shape != first.ensemble.shape:
                raise ValueError("DeviceSplit needs graphs of one shape (a static station graph)")
        self.x = torch.stack([g.x for g in graphs]).float().contiguous().to(device)
        self.ensemble = torch.stack([g.ensemble for g in graphs]).float().contiguous().to(device)
        self.y = torch.stack([g.y.reshape(-1) for g in graphs]).float().contiguous().to(device)
        self.edge_index, self.edge_attr = first.edge_index, first.edge_attr
        self.num_stations = first.x.shape[0]
        self._graphs = {}

    def __len__(self):
        return self.x.shape[0]

    def batched_graph(self, b: int):
        """The station graph of a batch of `b` dates on the device (PyG collation of the static graph; built once per size -
        the ragged last batch of every epoch has the same one)."""
        hit = self._graphs.get(b)
        if hit is None:
            from ..graph import build_station_graph, collate_static
            ei, ea = collate_static(self.edge_index, self.edge_attr, self.num_stations, b)
            hit = self._graphs[b] = build_station_graph(ei, ea, b * self.num_stations).to(self.x.device)
        return hit

    def epoch_batches(self, batch_size: int, generator=None, shuffle: bool = True):
        """Device int64 index tensors of `batch_size` dates each; like the reference's DataLoader (train.py:155) the last
        batch is ragged when the split does not divide (the engine steps it outside the captured graph)."""
        n = len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        return list(order.to(self.x.device).split(batch_size))
