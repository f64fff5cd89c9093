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

    def tiles(self, hidden: int):
        """(forward, backward) StationTiles for rows of `hidden` floats, or None when the tiled kernels do not
        apply (small graphs, unsupported width, a row gathering more rows than shared memory holds).  Built on
        the host on first use and kept with the graph."""
        if hidden < 128 or hidden % 128 or hidden > 512:
            return None
        cache = self.__dict__.setdefault("_tiles", {})
        if "pair" not in cache:                 # tiles hold 128-column chunks of rows: one layout for every width
            cache["pair"] = _build_tiles_pair(self, hidden)
        return cache["pair"]

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


TILE_ROW_BYTES = 512       # the kernels stage one 128-column fp32 chunk of a row at a time
TILED_MIN_ROWS = 16384     # below this the warp-per-row kernels win (too few tiles to fill 148 SMs)


class StationTiles:
    """Tile-major layout of one gather matrix for rc_gine_aggr_{fwd,bwd}_tiled (include/rc_b200.h: rc_gine_tiles)."""
    _ARRAYS = ("tile_stage_ptr", "tile_blk_ptr", "stage_id", "blocks")

    def __init__(self, n_tiles: int, max_staged: int, max_block_bytes: int, row_bytes: int, num_rows: int, arrays: dict,
                 n_entries: int = 0):
        self.n_tiles, self.max_staged, self.max_block_bytes, self.row_bytes = n_tiles, max_staged, max_block_bytes, row_bytes
        self.num_rows = num_rows
        self.n_entries = n_entries          # shared-memory row reads per pass (edges / n_entries = reuse factor)
        self.arrays = arrays
        # tile counters of the forward kernel's dynamic schedule (zero between launches)
        self.sched = torch.zeros(8, dtype=torch.int32, device=arrays["blocks"].device)
        self.struct = _lib.rc_gine_tiles(n_tiles, max_staged, max_block_bytes, row_bytes,
                                         *[arrays[a].data_ptr() for a in self._ARRAYS], self.sched.data_ptr())

    @property
    def n_staged(self) -> int:
        return int(self.arrays["stage_id"].numel())

    @property
    def n_halo(self) -> int:
        """Staged rows that their tile does not own (0 for a batch of reference graphs)."""
        return self.n_staged - self.num_rows

    def to(self, device) -> "StationTiles":
        return StationTiles(self.n_tiles, self.max_staged, self.max_block_bytes, self.row_bytes, self.num_rows,
                            {k: v.to(device) for k, v in self.arrays.items()}, self.n_entries)

    def verify(self, rowptr: torch.Tensor, col: torch.Tensor, attr: torch.Tensor) -> None:
        """rc_gine_tiles_verify_host: raises RcError unless the tiles hold exactly the edges of the CSR."""
        rowptr, col, attr = (t.detach().cpu().contiguous() for t in (rowptr, col, attr))
        a = {k: v.cpu().contiguous() for k, v in self.arrays.items()}
        e = col.numel()
        if e == 0:
            col, attr = torch.zeros(1, dtype=torch.int32), torch.zeros(1)
        _lib.check(_lib.lib().rc_gine_tiles_verify_host(
            rowptr.data_ptr(), col.data_ptr(), attr.data_ptr(), rowptr.numel() - 1, e, self.n_tiles, self.max_staged,
            self.max_block_bytes, self.row_bytes, *[a[k].data_ptr() for k in self._ARRAYS]), "rc_gine_tiles_verify_host")


def tile_limits(hidden: int):
    """(max_src, max_block_bytes) one CTA can hold at `hidden` columns, or None when the width is unsupported."""
    ms, mb = C.c_int(), C.c_int()
    if _lib.lib().rc_gine_tiles_limits(int(hidden), C.addressof(ms), C.addressof(mb)) != 0:
        return None
    return ms.value, mb.value


def build_tiles_host(rowptr: torch.Tensor, col: torch.Tensor, attr: torch.Tensor, max_src: int, max_block_bytes: int,
                     row_bytes: int) -> StationTiles:
    """rc_gine_tiles_build_host on host copies of a CSR (rowptr [M+1] int32, col [E] int32, attr [E] float32)."""
    L = _lib.lib()
    rowptr, col, attr = (t.detach().cpu().contiguous() for t in (rowptr, col, attr))
    n, e = rowptr.numel() - 1, col.numel()
    if e == 0:                                   # empty tensors have no storage address: give the C side one
        col, attr = torch.zeros(1, dtype=torch.int32), torch.zeros(1)
    tsp = torch.zeros(n + 1, dtype=torch.int32)
    tbp = torch.zeros(n + 1, dtype=torch.int32)
    stage = torch.empty(max(n + e, 1), dtype=torch.int32)
    blocks = torch.empty(max(16 * n + 4 * e, 4), dtype=torch.int32)
    nt, ms, mb = C.c_int32(), C.c_int32(), C.c_int32()
    ns, nu, ne = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(L.rc_gine_tiles_build_host(rowptr.data_ptr(), col.data_ptr(), attr.data_ptr(), n, e, int(max_src),
                                          int(max_block_bytes), int(row_bytes), tsp.data_ptr(), tbp.data_ptr(),
                                          stage.data_ptr(), blocks.data_ptr(), C.addressof(nt), C.addressof(ns),
                                          C.addressof(nu), C.addressof(ms), C.addressof(mb), C.addressof(ne)),
               "rc_gine_tiles_build_host")
    arrays = {"tile_stage_ptr": tsp[:nt.value + 1].clone(), "tile_blk_ptr": tbp[:nt.value + 1].clone(),
              "stage_id": stage[:ns.value].clone() if ns.value else torch.zeros(1, dtype=torch.int32)[:0],
              "blocks": blocks[:max(4 * nu.value, 4)].clone()}
    return StationTiles(nt.value, ms.value, mb.value, int(row_bytes), n, arrays, ne.value)


def _build_tiles_pair(g: "StationGraph", hidden: int):
    lim = tile_limits(hidden)
    if lim is None or g.num_nodes < TILED_MIN_ROWS:
        return None
    try:
        fwd = build_tiles_host(g.rowptr, g.col, g.attr, lim[0], lim[1], TILE_ROW_BYTES)
        bwd = build_tiles_host(g.t_rowptr, g.t_dst, g.t_attr, lim[0], lim[1], TILE_ROW_BYTES)
    except _lib.RcError:
        return None            # some row gathers more rows than one CTA can stage: the untiled kernels handle it
    return fwd.to(g.device), bwd.to(g.device)


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
