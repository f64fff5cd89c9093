"""Oracle (test infrastructure): station-graph construction and layouts, numpy only.

Restates
  * utils/data.py:261-284  build_edge_index_and_attr  (radius graph + self loops)
  * PyG Batch.from_data_list edge_index collation [PyG-upstream, not vendored]:
    `edge_index` of graph i is shifted by i*N and concatenated on dim 1
    (call sites train.py:155-156, eval.py:141)
  * the dst-sorted CSR / transpose / reverse-edge layout the CUDA path must
    reproduce bit-exactly.
Integer outputs are exact; `edge_attr` is float32 computed with the same numpy
expression as the reference so that it is bitwise identical.
"""
from __future__ import annotations

import numpy as np


def radius_graph(dist_mat: np.ndarray, max_dist: float):
    """utils/data.py:261-284.  Returns (edge_index int64 [2,E], edge_attr f32 [E,1]).

    Non-self edges come out of np.where in row-major order (src ascending, then
    dst ascending; src = edge_index[0]); N self loops with attr 1.0 are appended.
    """
    d = np.array(dist_mat, copy=True)
    n = d.shape[0]
    d[np.arange(n), np.arange(n)] = np.inf                      # :266
    src, dst = np.nonzero(d <= max_dist)                        # :267 (row-major)
    vals = d[src, dst]                                          # :268
    top = vals.max() if vals.size > 0 else 1.0                  # :269
    inv = (vals / top) ** -1                                    # :272
    loops = np.arange(n, dtype=np.int64)
    edge_index = np.stack([np.concatenate([src.astype(np.int64), loops]),
                           np.concatenate([dst.astype(np.int64), loops])])
    edge_attr = np.concatenate([inv.astype(np.float32), np.ones(n, np.float32)])[:, None]
    return edge_index, edge_attr


def collate_edges(edge_index: np.ndarray, edge_attr: np.ndarray, num_nodes: int, batch: int):
    """PyG Batch.from_data_list for a static graph repeated `batch` times."""
    e = edge_index.shape[1]
    off = (np.arange(batch, dtype=np.int64) * num_nodes).repeat(e)
    ei = np.tile(edge_index, (1, batch)) + off[None, :]
    ea = np.tile(edge_attr, (batch, 1))
    return ei, ea


def collate_edge_list(graphs):
    """General collate: graphs = [(edge_index, edge_attr, num_nodes), ...]."""
    eis, eas, off = [], [], 0
    for ei, ea, n in graphs:
        eis.append(ei + off)
        eas.append(ea)
        off += n
    return np.concatenate(eis, axis=1), np.concatenate(eas, axis=0), off


def csr_layout(edge_index: np.ndarray, edge_attr: np.ndarray, num_nodes: int) -> dict:
    """dst-sorted CSR + src-sorted transpose + reverse-edge map (all int32 / f32).

    rowptr[M+1]; for slot s in row i (rowptr[i] <= s < rowptr[i+1]):
      col[s]  = source node of the edge,   attr[s] = its edge_attr (bitwise),
      perm[s] = index of that edge in the reference `edge_index` (stable sort by dst).
    Transpose (stable sort by src): t_rowptr[M+1]; for position q in row j:
      t_dst[q] = destination, t_attr[q] = attr, t_perm[q] = reference edge id,
      t_slot[q] = the edge's slot in the dst-sorted order.
    rev[s] = slot of the edge (dst->src) for slot s = (src->dst), or -1.
    """
    src = edge_index[0].astype(np.int64)
    dst = edge_index[1].astype(np.int64)
    e = src.shape[0]
    perm = np.argsort(dst, kind="stable")
    rowptr = np.zeros(num_nodes + 1, np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    t_perm = np.argsort(src, kind="stable")
    t_rowptr = np.zeros(num_nodes + 1, np.int64)
    np.add.at(t_rowptr, src + 1, 1)
    t_rowptr = np.cumsum(t_rowptr)
    inv = np.empty(e, np.int64)
    inv[perm] = np.arange(e)
    # reverse-edge map via a dictionary on (src, dst) -> first slot
    key = src[perm] * num_nodes + dst[perm]          # key of slot s
    order = np.argsort(key, kind="stable")
    skey = key[order]
    want = dst[perm] * num_nodes + src[perm]
    pos = np.searchsorted(skey, want, side="left")
    pos_c = np.minimum(pos, max(e - 1, 0))
    found = (pos < e) & (skey[pos_c] == want) if e > 0 else np.zeros(0, bool)
    rev = np.where(found, order[pos_c], -1) if e > 0 else np.zeros(0, np.int64)
    flat_attr = edge_attr.reshape(-1)
    return {
        "rowptr": rowptr.astype(np.int32),
        "col": src[perm].astype(np.int32),
        "attr": flat_attr[perm].astype(np.float32),
        "perm": perm.astype(np.int32),
        "t_rowptr": t_rowptr.astype(np.int32),
        "t_dst": dst[t_perm].astype(np.int32),
        "t_attr": flat_attr[t_perm].astype(np.float32),
        "t_perm": t_perm.astype(np.int32),
        "t_slot": inv[t_perm].astype(np.int32),
        "rev": rev.astype(np.int32),
    }


def synthetic_coords(num_nodes: int, box: float, seed: int = 0) -> np.ndarray:
    """default_rng(seed).uniform(0, box, (N, 2)) as km coordinates."""
    return np.random.default_rng(seed).uniform(0.0, box, (num_nodes, 2))


def euclid_dist_matrix(coords: np.ndarray) -> np.ndarray:
    """float32 Euclidean distance matrix (stands in for compute_dist_matrix, utils/data.py:248-259)."""
    diff = coords[:, None, :] - coords[None, :, :]
    return np.sqrt((diff * diff).sum(-1)).astype(np.float32)
