"""Station-graph layout handed to the GINE kernels: dst-sorted CSR, src-sorted transpose and the
reverse-edge map, built once per distinct (graph, batch size) and cached (the reference's graph is
static: every date shares one edge_index / edge_attr, utils/data.py:300,330-335).

Replaces PyG's per-step collate of `edge_index` (train.py:155-156 -> Batch.from_data_list) and the
index_select / scatter indices of GINEConv.  All arithmetic is in librc_b200.so (rc_csr_build_host /
rc_csr_build, rc_radius_graph_*); this module only owns the buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

_FIELDS = ("rowptr", "col", "attr", "perm", "t_rowptr", "t_dst", "t_attr", "t_perm", "t_slot", "rev")
_FLOAT = {"attr", "t_attr"}


@dataclass
class StationGraph:
    """CSR / transpose layout of one batched station graph (int32 indices, float32 attributes)."""
    num_nodes: int
    num_edges: int
    rowptr: torch.Tensor
    col: torch.Tensor
    attr: torch.Tensor
    perm: torch.Tensor
    t_rowptr: torch.Tensor
    t_dst: torch.Tensor
    t_attr: torch.Tensor
    t_perm: torch.Tensor
    t_slot: torch.Tensor
    rev: torch.Tensor

    @property
    def device(self):
        return self.rowptr.device

    def to(self, device, non_blocking: bool = False) -> "StationGraph":
        if torch.device(device) == self.device:
            return self
        return StationGraph(self.num_nodes, self.num_edges,
                            *[getattr(self, f).to(device, non_blocking=non_blocking) for f in _FIELDS])

    def edge_index(self) -> torch.Tensor:
        """Reconstruct the reference `edge_index` (int64 [2,E]) — the bit-exactness check of Appendix B."""
        rows = torch.repeat_interleave(torch.arange(self.num_nodes, device=self.device),
                                       (self.rowptr[1:] - self.rowptr[:-1]).long())
        inv = torch.empty(self.num_edges, dtype=torch.long, device=self.device)
        inv[self.perm.long()] = torch.arange(self.num_edges, device=self.device)
        return torch.stack([self.col.long(), rows])[:, inv]


def _alloc(num_nodes: int, num_edges: int, device) -> dict:
    out = {}
    for f in _FIELDS:
        n = num_nodes + 1 if f.endswith("rowptr") else num_edges
        # at least one element of storage so that an empty graph still has non-null pointers
        out[f] = torch.empty(max(n, 1), dtype=torch.float32 if f in _FLOAT else torch.int32, device=device)[:n]
    return out


def _csr_struct(bufs: dict) -> _lib.rc_csr:
    return _lib.rc_csr(**{f: bufs[f].data_ptr() for f in _FIELDS})


def build_station_graph(edge_index: torch.Tensor, edge_attr: torch.Tensor, num_nodes: int) -> StationGraph:
    """CSR layout of `edge_index` (int64 [2,E]) / `edge_attr` ([E] or [E,1]); host tensors are laid out by
    rc_csr_build_host, CUDA tensors by the rc_csr_build kernels — same result bit for bit."""
    L = _lib.lib()
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError(f"edge_index must be [2, E], got {tuple(edge_index.shape)}")
    ei = edge_index.to(torch.int64).contiguous()
    ea = edge_attr.reshape(-1).to(torch.float32).contiguous()
    e = int(ei.shape[1])
    if ea.numel() != e:
        raise ValueError(f"edge_attr has {ea.numel()} entries for {e} edges")
    bufs = _alloc(num_nodes, e, ei.device)
    st = _csr_struct(bufs)
    if ei.is_cuda:
        ws_bytes = int(L.rc_csr_build_workspace(e, num_nodes))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ei.device)
        flag = torch.zeros(1, dtype=torch.int32, device=ei.device)
        with torch.cuda.device(ei.device):
            _lib.check(L.rc_csr_build(ei.data_ptr(), ea.data_ptr(), e, num_nodes, C.byref(st), ws.data_ptr(), ws_bytes,
                                      flag.data_ptr(), _lib.stream_ptr(ei.device)), "rc_csr_build")
        if not torch.cuda.is_current_stream_capturing() and int(flag.item()) != 0:
            raise _lib.RcError("rc_csr_build: edge_index holds a node id outside [0, num_nodes)")
    else:
        _lib.check(L.rc_csr_build_host(ei.data_ptr(), ea.data_ptr(), e, num_nodes, C.byref(st)), "rc_csr_build_host")
    return StationGraph(num_nodes, e, **bufs)


def radius_graph(dist_mat, max_dist: float):
    """utils/data.py:261-284 (build_edge_index_and_attr): returns (edge_index int64 [2,E], edge_attr f32 [E,1])."""
    L = _lib.lib()
    d = np.ascontiguousarray(np.asarray(dist_mat), dtype=np.float32)
    if d.ndim != 2 or d.shape[0] != d.shape[1]:
        raise ValueError("dist_mat must be square")
    n = d.shape[0]
    cnt = C.c_int64()
    _lib.check(L.rc_radius_graph_count_host(d.ctypes.data, n, float(max_dist), C.addressof(cnt)), "rc_radius_graph_count_host")
    ei = torch.empty((2, cnt.value), dtype=torch.int64)
    ea = torch.empty((cnt.value, 1), dtype=torch.float32)
    if cnt.value > 0:
        _lib.check(L.rc_radius_graph_fill_host(d.ctypes.data, n, float(max_dist), cnt.value, ei.data_ptr(), ea.data_ptr()),
                   "rc_radius_graph_fill_host")
    return ei, ea


def radius_graph_from_coords(coords, max_dist: float):
    """Same edge order and attributes as `radius_graph(distance_matrix(coords))` without the N x N matrix."""
    L = _lib.lib()
    xy = np.ascontiguousarray(np.asarray(coords), dtype=np.float64)
    if xy.ndim != 2 or xy.shape[1] != 2:
        raise ValueError("coords must be [N, 2]")
    n = xy.shape[0]
    cnt = C.c_int64()
    _lib.check(L.rc_radius_graph_coords_count_host(xy.ctypes.data, n, float(max_dist), C.addressof(cnt)),
               "rc_radius_graph_coords_count_host")
    ei = torch.empty((2, cnt.value), dtype=torch.int64)
    ea = torch.empty((cnt.value, 1), dtype=torch.float32)
    if cnt.value > 0:
        _lib.check(L.rc_radius_graph_coords_fill_host(xy.ctypes.data, n, float(max_dist), cnt.value, ei.data_ptr(), ea.data_ptr()),
                   "rc_radius_graph_coords_fill_host")
    return ei, ea


def collate_static(edge_index: torch.Tensor, edge_attr: torch.Tensor, num_nodes: int, batch: int):
    """PyG Batch.from_data_list for one static graph repeated `batch` times: edge_index of copy i shifted by i*N."""
    e = edge_index.shape[1]
    off = (torch.arange(batch, dtype=torch.int64) * num_nodes).repeat_interleave(e)
    return edge_index.repeat(1, batch) + off.unsqueeze(0), edge_attr.reshape(-1, 1).repeat(batch, 1)


class GraphCache:
    """(edge tensors identity, num_nodes) -> StationGraph.  A batch that arrives with plain PyG-style
    `edge_index` / `edge_attr` (train.py:62-64) is laid out on first sight and reused afterwards."""

    def __init__(self, capacity: int = 16):
        self.capacity = capacity
        self._items: dict = {}

    @staticmethod
    def _key(edge_index, edge_attr, num_nodes):
        return (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, edge_attr.data_ptr(),
                edge_attr._version, str(edge_index.device), num_nodes)

    def get(self, edge_index, edge_attr, num_nodes) -> StationGraph:
        key = self._key(edge_index, edge_attr, num_nodes)
        hit = self._items.get(key)
        if hit is not None:
            return hit[0]
        g = build_station_graph(edge_index, edge_attr, num_nodes)
        if len(self._items) >= self.capacity:
            self._items.pop(next(iter(self._items)))
        self._items[key] = (g, edge_index, edge_attr)      # keep the tensors alive so the pointers stay unique
        return g


def graph_of(data) -> StationGraph:
    """The StationGraph of a batch: the loader-attached one (`data.station_graph`) or a cached build."""
    g = getattr(data, "station_graph", None)
    dev = data.x.device
    if g is not None:
        if g.device != dev:
            g = g.to(dev)
            data.station_graph = g
        return g
    return _GLOBAL_CACHE.get(data.edge_index, data.edge_attr, int(data.x.shape[0]))


_GLOBAL_CACHE = GraphCache()
