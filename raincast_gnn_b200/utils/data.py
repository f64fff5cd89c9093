"""Mirror of the graph helpers of utils/data.py that sit on the hot path.  The Zarr/pandas ingestion
(utils/data.py:19-237,347-411) is offline preprocessing and out of scope."""
from __future__ import annotations

from typing import List

import torch

from ..graph import radius_graph
from ..pyg_compat.data import Data


def build_edge_index_and_attr(dist_mat, max_dist: float):
    """utils/data.py:261-284, computed by rc_radius_graph_*_host; bit-identical edge order and attributes."""
    return radius_graph(dist_mat, max_dist)


def make_graphs(x: torch.Tensor, ensemble: torch.Tensor, y: torch.Tensor, edge_index: torch.Tensor,
                edge_attr: torch.Tensor, num_stations: int, timestamps=None) -> List[Data]:
    """One Data per date sharing ONE edge_index / edge_attr object (utils/data.py:300,330-337)."""
    n_dates = x.shape[0] // num_stations
    graphs = []
    for i in range(n_dates):
        s = slice(i * num_stations, (i + 1) * num_stations)
        d = Data(x=x[s], ensemble=ensemble[s], edge_index=edge_index, edge_attr=edge_attr, y=y[s])
        d.timestamp = timestamps[i] if timestamps is not None else i
        graphs.append(d)
    return graphs


def split_graph(graph: Data, new_gnn: bool = False) -> List[Data]:
    """utils/data.py:418-446 (new_gnn=True branch, the one eval.py:134 uses): 51 members -> 5 graphs of 10."""
    if not new_gnn:
        raise NotImplementedError("the node-permutation branch (utils/data.py:432-446) is unused by eval.py")
    bounds = [0, 10, 20, 30, 40, 50]
    out = []
    for i in range(5):
        g = graph.clone()
        g.edge_index, g.edge_attr = graph.edge_index, graph.edge_attr     # keep the shared static graph
        g.ensemble = g.ensemble[:, bounds[i]:bounds[i + 1], :]
        out.append(g)
    return out
