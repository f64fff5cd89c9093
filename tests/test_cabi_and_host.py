"""CPU-side checks: the C-ABI library loads and exports every symbol include/rc_b200.h declares; the
host-side graph builders are bit-exact against the reference fixtures and the oracle; the loader seam
behaves like PyG's collate.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import graph as og
from raincast_gnn_b200 import _lib, graph as G
from raincast_gnn_b200.utils import synthetic as syn


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rc_b200.h")).read()
    declared = set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in rc_b200.h but not exported"
    _lib.lib()
    assert set(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ declared
    assert _lib.lib().rc_version() == 100


def test_struct_sizes_match_header(tmp_path):
    """ctypes mirrors must have the C layout: sizes and a few field offsets as gcc lays out include/rc_b200.h."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rc_b200.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(rc_csr), sizeof(rc_operand), sizeof(rc_reduce_seg), sizeof(rc_gemm), sizeof(rc_gine_tiles), '
                   'offsetof(rc_gemm, stats), offsetof(rc_gemm, tc_ws), offsetof(rc_gemm, a_out)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    want = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    got = [ctypes.sizeof(_lib.rc_csr), ctypes.sizeof(_lib.rc_operand), ctypes.sizeof(_lib.rc_reduce_seg), ctypes.sizeof(_lib.rc_gemm),
           ctypes.sizeof(_lib.rc_gine_tiles), _lib.rc_gemm.stats.offset, _lib.rc_gemm.tc_ws.offset, _lib.rc_gemm.a_out.offset]
    assert got == want


def test_errors_are_reported_not_swallowed():
    L = _lib.lib()
    assert L.rc_gemm_run(None, None) == 1
    assert b"rc_gemm_run" in L.rc_last_error()
    with pytest.raises(_lib.RcError):
        _lib.check(L.rc_crps_fwd_bwd(None, None, None, None, None, 4, 3, 0, 0.0, 0.5, 5.0, None, 0, None), "rc_crps_fwd_bwd")
    with pytest.raises(_lib.RcError):        # CPU tensors never reach a kernel
        _lib.require_cuda(torch.zeros(3))


@pytest.mark.parametrize("name,n,box,md", [("ref122_d100", 122, 600.0, 100.0), ("ref122_d1", 122, 600.0, 1.0),
                                           ("n7_d300", 7, 600.0, 300.0), ("n40_d150", 40, 600.0, 150.0)])
def test_radius_graph_host_bit_exact(golden_graph, name, n, box, md):
    coords = syn.station_coords(n, box, seed=0)
    for ei, ea in (G.radius_graph(syn.distance_matrix(coords), md), G.radius_graph_from_coords(coords, md)):
        assert ei.dtype == torch.int64 and ea.dtype == torch.float32 and ea.shape == (ei.shape[1], 1)
        assert np.array_equal(ei.numpy(), golden_graph[f"{name}.edge_index"])
        assert np.array_equal(ea.numpy().view(np.uint32), golden_graph[f"{name}.edge_attr"].view(np.uint32))


def test_radius_graph_asymmetric_and_empty(golden_graph):
    ei, ea = G.radius_graph(golden_graph["asym23.dist"], 120.0)
    assert np.array_equal(ei.numpy(), golden_graph["asym23.edge_index"])
    assert np.array_equal(ea.numpy(), golden_graph["asym23.edge_attr"])
    ei, ea = G.radius_graph(np.zeros((0, 0), np.float32), 1.0)
    assert ei.shape == (2, 0) and ea.shape == (0, 1)


@pytest.mark.parametrize("name,batch", [("ref122_d100", 8), ("asym23", 3), ("ref122_d1", 2)])
def test_csr_host_matches_oracle(golden_graph, name, batch):
    ei, ea = torch.from_numpy(golden_graph[f"{name}.edge_index"]), torch.from_numpy(golden_graph[f"{name}.edge_attr"])
    n = int(ei.max()) + 1
    eib, eab = G.collate_static(ei, ea, n, batch)
    oei, oea = og.collate_edges(ei.numpy(), ea.numpy(), n, batch)
    assert np.array_equal(eib.numpy(), oei) and np.array_equal(eab.numpy(), oea)
    sg = G.build_station_graph(eib, eab, n * batch)
    for k, v in og.csr_layout(oei, oea, n * batch).items():
        assert np.array_equal(getattr(sg, k).numpy().view(np.uint32), v.view(np.uint32)), k
    assert torch.equal(sg.edge_index(), eib)


def test_scaled_graph_edge_count():
    n = 100_000
    ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
    assert ei.shape[1] == 2_978_560          # SURVEY.md 8d [probe]
    assert float(ea.min()) == 1.0


def test_coords_graph_matches_kdtree():
    from scipy.spatial import cKDTree
    n, r = 3000, 25.0
    coords = syn.station_coords(n, 1000.0, seed=5)
    ei, _ = G.radius_graph_from_coords(coords, r)
    pairs = cKDTree(coords).query_pairs(r * 1.0001, output_type="ndarray")
    d = np.sqrt(((coords[pairs[:, 0]] - coords[pairs[:, 1]]) ** 2).sum(-1)).astype(np.float32)
    pairs = pairs[d <= np.float32(r)]
    want = set(map(tuple, pairs)) | set((b, a) for a, b in pairs)
    got = set(map(tuple, ei[:, :-n].T.tolist()))
    assert got == want
    assert torch.equal(ei[:, -n:], torch.arange(n).repeat(2, 1))


def test_loader_collates_like_pyg_and_caches_the_graph():
    from oracle import pyg as opyg
    from raincast_gnn_b200.pyg_compat import Batch, DataLoader
    from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
    ds = SyntheticEUPPBench(n_dates=10, num_stations=20, members=3, feats=4, max_dist=250.0)
    loader = DataLoader(ds, batch_size=4, shuffle=False)
    assert len(loader) == 3
    batches = list(loader)
    ref = opyg.Batch.from_data_list([opyg.Data(x=d.x, ensemble=d.ensemble, edge_index=d.edge_index, edge_attr=d.edge_attr, y=d.y)
                                     for d in ds.graphs[:4]])
    b0 = batches[0]
    for k in ("x", "ensemble", "edge_index", "edge_attr", "batch", "ptr"):
        assert torch.equal(getattr(b0, k), getattr(ref, k)), k
    assert torch.equal(torch.nan_to_num(b0.y, nan=-99.0), torch.nan_to_num(ref.y, nan=-99.0))
    assert batches[1].station_graph is b0.station_graph            # static graph: one CSR per batch size
    assert batches[2].station_graph is not b0.station_graph and batches[2].x.shape[0] == 40
    assert torch.equal(b0.station_graph.edge_index(), b0.edge_index)


def test_split_graph_and_compat_install():
    from raincast_gnn_b200 import pyg_compat
    from raincast_gnn_b200.utils.data import split_graph
    from raincast_gnn_b200.utils.dataset import EUPPBench, SyntheticEUPPBench
    ds = SyntheticEUPPBench(n_dates=1, members=51)
    parts = split_graph(ds[0], True)
    assert len(parts) == 5 and all(p.ensemble.shape[1] == 10 for p in parts)
    assert torch.equal(parts[2].ensemble, ds[0].ensemble[:, 20:30])
    assert parts[0].edge_index is ds[0].edge_index
    with pytest.raises(ValueError):
        EUPPBench("raw", "processed", split="bogus")
    with pytest.raises(FileNotFoundError):
        EUPPBench("raw", "processed", split="train_rf")
    pyg_compat.install()
    from torch_geometric.loader import DataLoader  # noqa: F401


def test_model_state_dict_layout_and_ctor_errors(golden_model):
    from raincast_gnn_b200.models import GNN, MixedLoss, ResGnn
    m = GNN(35, 128, 128, 4, torch.optim.AdamW, {"lr": 1e-4}, "MixedLoss", "True", 1.71, 0.5)
    assert list(m.state_dict().keys()) == list(golden_model["ref_mixed_u.keys"])
    assert sum(p.numel() for p in m.parameters()) == 209_929
    assert m.out_channels == 5 and isinstance(m.loss_fn, MixedLoss) and m.loss_fn.grad_u is True
    assert GNN(35, 128, 128, 4, None, None, "MixedLoss", "False", 1.71, 0.5).out_channels == 4
    assert GNN(35, 128, 128, 4, None, None, "MixedNormalCRPS").out_channels == 3
    assert GNN(35, 128, 128, 4, None, None, "NormalCRPS").out_channels == 2
    with pytest.raises(AssertionError):
        ResGnn(8, 8, 0, 8)                                            # models/gnn.py:13
    with pytest.raises(ValueError):
        MixedLoss(grad_u=False, xi=1.0, u=1.0)


# ------------------------------------------------------------------------------------------------ station tiles
def _check_tiles(sg_rowptr, sg_col, sg_attr, tiles, max_src, max_block, row_bytes=512):
    """Invariants of rc_gine_tiles_build_host, decoded here independently of rc_gine_tiles_verify_host: a partition of
    the rows into groups of at most three, at most max_src staged rows and max_block record bytes per tile, and every
    row's edges recoverable from its group's entries with bit-identical attributes."""
    tiles.verify(torch.from_numpy(sg_rowptr), torch.from_numpy(sg_col), torch.from_numpy(sg_attr))
    a = {k: v.numpy() for k, v in tiles.arrays.items()}
    n = len(sg_rowptr) - 1
    tsp, tbp, stage = a["tile_stage_ptr"], a["tile_blk_ptr"], a["stage_id"]
    blocks = a["blocks"].reshape(-1, 4)
    assert len(tsp) == tiles.n_tiles + 1 and tsp[0] == 0 and tsp[-1] == len(stage) and tbp[0] == 0
    owned, staged_max, blk_max, edges, entries = [], 0, 0, 0, 0
    for t in range(tiles.n_tiles):
        blk = blocks[tbp[t]:tbp[t + 1]]
        nrows, nst, ngroups, ne = blk[0]
        staged = stage[tsp[t]:tsp[t + 1]]
        assert nst == len(staged) <= max_src and len(np.unique(staged)) == nst and 16 * len(blk) <= max_block
        staged_max, blk_max, edges = max(staged_max, nst), max(blk_max, 16 * len(blk)), edges + ne
        r = 0
        for g in range(ngroups):
            u0, u1, u2 = blk[1 + 3 * g: 4 + 3 * g]
            cnt = [0, u1[0] & 0xffff, (u1[0] >> 16) & 0xffff, u1[1] & 0xffff, (u1[1] >> 16) & 0xffff, u1[2] & 0xffff,
                   (u1[2] >> 16) & 0xffff, u1[3]]
            assert u0[3] % 16 == 0 and u2[3] == r * row_bytes       # own rows lead the stage list, group by group
            ent = blk[u0[3] // 16: u0[3] // 16 + sum(cnt)]
            entries += len(ent)
            cls = np.repeat(np.arange(8), cnt)
            assert (ent[:, 0] % row_bytes == 0).all()
            for k in range(3):
                v = u0[k]
                if v < 0:
                    assert (u0[k:3] < 0).all() and not (cls & (1 << k)).any()
                    continue
                assert v == staged[r] and u2[k:k + 1].view(np.float32)[0] == sg_rowptr[v + 1] - sg_rowptr[v]
                r += 1
                owned.append(v)
                use = (cls & (1 << k)) != 0
                got = sorted(zip(staged[ent[use, 0] // row_bytes].tolist(), ent[use, 1 + k].view(np.uint32).tolist()))
                want = sorted(zip(sg_col[sg_rowptr[v]:sg_rowptr[v + 1]].tolist(),
                                  sg_attr[sg_rowptr[v]:sg_rowptr[v + 1]].view(np.uint32).tolist()))
                assert got == want
        assert r == nrows
    assert np.array_equal(np.sort(np.array(owned, dtype=np.int64)), np.arange(n))
    assert edges == len(sg_col) and entries == tiles.n_entries
    assert staged_max == tiles.max_staged and blk_max == tiles.max_block_bytes


def test_station_tiles_batched_reference_graph(golden_graph):
    """A batch of reference graphs is block-diagonal: one tile per 122-station graph, no halo."""
    ei, ea = golden_graph["ref122_d100.edge_index"], golden_graph["ref122_d100.edge_attr"]
    batch = 12
    ei_b, ea_b = og.collate_edges(ei, ea, 122, batch)
    sg = G.build_station_graph(torch.from_numpy(ei_b), torch.from_numpy(ea_b), 122 * batch)
    max_src, max_block = G.tile_limits(128)
    assert 2 * (max_src * 512 + max_block) + 64 <= 227 * 1024 and max_src >= 160      # two buffers + control block
    for rp, col, attr in ((sg.rowptr, sg.col, sg.attr), (sg.t_rowptr, sg.t_dst, sg.t_attr)):
        tiles = G.build_tiles_host(rp, col, attr, max_src, max_block, 512)
        _check_tiles(rp.numpy(), col.numpy(), attr.numpy(), tiles, max_src, max_block)
        assert tiles.n_tiles == batch and tiles.n_halo == 0 and tiles.max_staged == 122


@pytest.mark.parametrize("n,deg,max_src,max_block", [(3000, 12.0, 64, 1 << 20), (5000, 29.0, 175, 25584), (400, 6.0, 40, 1024),
                                                     (2000, 20.0, 200, 2048)])
def test_station_tiles_radius_graph(n, deg, max_src, max_block):
    coords = syn.station_coords(n, 300.0, seed=3)
    ei, ea = G.radius_graph_from_coords(coords, syn.scaled_graph_radius(n, 300.0, deg))
    sg = G.build_station_graph(ei, ea, n)
    tiles = G.build_tiles_host(sg.rowptr, sg.col, sg.attr, max_src, max_block, 512)
    _check_tiles(sg.rowptr.numpy(), sg.col.numpy(), sg.attr.numpy(), tiles, max_src, max_block)
    # clusters are compact: each owned row brings well under its full neighbourhood in halo rows
    assert tiles.n_halo < 0.5 * sg.num_edges


def test_station_tiles_edge_cases():
    # asymmetric graph with an isolated node, a multi-edge and no self loops
    ei = torch.tensor([[0, 0, 1, 3, 3, 4], [1, 1, 2, 0, 2, 0]])
    ea = torch.arange(6, dtype=torch.float32) + 1.0
    sg = G.build_station_graph(ei, ea, 6)
    for rp, col, attr in ((sg.rowptr, sg.col, sg.attr), (sg.t_rowptr, sg.t_dst, sg.t_attr)):
        for max_src in (3, 4, 100):
            tiles = G.build_tiles_host(rp, col, attr, max_src, 4096, 1024)
            _check_tiles(rp.numpy(), col.numpy(), attr.numpy(), tiles, max_src, 4096, row_bytes=1024)
    # a row that cannot fit: node 0 gathers {3, 4} plus itself = 3 staged rows > 2; or its records exceed the block
    with pytest.raises(_lib.RcError):
        G.build_tiles_host(sg.rowptr, sg.col, sg.attr, 2, 4096, 512)
    star = G.build_station_graph(torch.tensor([[1, 2, 3, 4, 5], [0, 0, 0, 0, 0]]), torch.ones(5), 6)
    # the hub's group: header 16 + group record 48 + 6 entries (5 leaves + the hub's own row) 96 = 160 bytes
    with pytest.raises(_lib.RcError):
        G.build_tiles_host(star.rowptr, star.col, star.attr, 100, 128, 512)
    _check_tiles(star.rowptr.numpy(), star.col.numpy(), star.attr.numpy(),
                 G.build_tiles_host(star.rowptr, star.col, star.attr, 100, 256, 512), 100, 256)
    # empty graph
    empty = G.build_station_graph(torch.zeros((2, 0), dtype=torch.int64), torch.zeros(0), 0)
    tiles = G.build_tiles_host(empty.rowptr, empty.col, empty.attr, 8, 4096, 512)
    assert tiles.n_tiles == 0 and tiles.n_staged == 0
    assert G.tile_limits(100) is None and G.tile_limits(512) == G.tile_limits(128)


def test_station_tiles_verifier_rejects_corruption():
    """rc_gine_tiles_verify_host is the CPU check of the layout the tiled kernels walk: it must notice a wrong
    attribute, a wrong staged offset, a row owned twice and a missing edge."""
    coords = syn.station_coords(600, 120.0, seed=5)
    ei, ea = G.radius_graph_from_coords(coords, syn.scaled_graph_radius(600, 120.0, 10.0))
    sg = G.build_station_graph(ei, ea, 600)
    tiles = G.build_tiles_host(sg.rowptr, sg.col, sg.attr, 64, 8192, 512)
    tiles.verify(sg.rowptr, sg.col, sg.attr)
    blocks = tiles.arrays["blocks"]
    tbp = tiles.arrays["tile_blk_ptr"].numpy()
    hdr = blocks[4 * tbp[0]: 4 * tbp[0] + 4].numpy()
    ent0 = 4 * tbp[0] + (16 + 48 * int(hdr[2])) // 4            # first entry of the first tile

    def corrupted(idx, value):
        saved = int(blocks[idx])
        blocks[idx] = value
        try:
            with pytest.raises(_lib.RcError):
                tiles.verify(sg.rowptr, sg.col, sg.attr)
        finally:
            blocks[idx] = saved
    rec = blocks[4 * tbp[0] + 4: 4 * tbp[0] + 16].numpy()      # group 0: {nodes, entry offset}{class counts}{degrees, self}
    cnt = [0, rec[4] & 0xffff, (rec[4] >> 16) & 0xffff, rec[5] & 0xffff, (rec[5] >> 16) & 0xffff, rec[6] & 0xffff,
           (rec[6] >> 16) & 0xffff, rec[7]]
    cls = max(c for c in range(1, 8) if cnt[c])                 # class of group 0's last entry
    last = ent0 + 4 * (sum(cnt) - 1)
    k = [b for b in range(3) if cls & (1 << b)][0]
    corrupted(last + 1 + k, int(blocks[last + 1 + k]) ^ 1)      # one attribute bit of a row that uses the entry
    corrupted(ent0, int(blocks[ent0]) + 512 * 4096)                                   # staged offset out of range
    corrupted(4 * tbp[0] + 4, int(blocks[4 * tbp[0] + 8]) if False else int(blocks[4 * tbp[0] + 5]))   # row 0 := row 1 (owned twice)
    corrupted(4 * tbp[0] + 8, 0)                                                      # class counts zeroed: edges missing
    tiles.verify(sg.rowptr, sg.col, sg.attr)                                          # restored


def test_device_split_epoch_batches():
    """DeviceSplit (utils/dataset.py) on the CPU: stacking, batches like the reference's DataLoader (ragged last batch kept,
    train.py:155), reproducible order."""
    from raincast_gnn_b200.utils.dataset import DeviceSplit, SyntheticEUPPBench
    ds = SyntheticEUPPBench(n_dates=11, num_stations=7, members=3, feats=5)
    split = DeviceSplit([ds[i] for i in range(len(ds))], "cpu")
    assert split.x.shape == (11, 7, 5) and split.ensemble.shape == (11, 7, 3, 5) and split.y.shape == (11, 7) and len(split) == 11
    assert torch.equal(split.x[4], ds[4].x) and torch.equal(split.ensemble[9], ds[9].ensemble)
    b1 = split.epoch_batches(4, generator=torch.Generator().manual_seed(3))
    b2 = split.epoch_batches(4, generator=torch.Generator().manual_seed(3))
    assert [b.numel() for b in b1] == [4, 4, 3] and all(b.dtype == torch.int64 for b in b1)
    assert all(torch.equal(a, b) for a, b in zip(b1, b2))
    assert sorted(torch.cat(b1).tolist()) == list(range(11))
    assert torch.equal(torch.cat(split.epoch_batches(5, shuffle=False)), torch.arange(11))
    with pytest.raises(ValueError):
        DeviceSplit([], "cpu")


def test_euppbench_reads_reference_processed_file_without_pyg(tmp_path):
    """utils/dataset.py:174-182 writes torch.save((data, slices)) with `data` a pickled torch_geometric Data (its
    attributes live in data._store._mapping, PyG 2.x) and utils/dataset.py:57-60 reads it back.  torch-geometric is not
    installable here: EUPPBench must read such a file with the PyG classes replaced by attribute bags.  The file is
    written here through stand-in classes registered under PyG's module paths, which are removed again before loading."""
    import sys
    import types
    from raincast_gnn_b200.utils.dataset import EUPPBench, SyntheticEUPPBench
    ds = SyntheticEUPPBench(n_dates=5, num_stations=17, members=3, feats=6, max_dist=250.0)
    mods = {}
    for name in ("torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.data.storage"):
        mods[name] = types.ModuleType(name)

    class GlobalStorage:                                   # torch_geometric.data.storage.GlobalStorage: attributes in _mapping
        def __init__(self, mapping, parent):
            self._mapping, self._parent = mapping, parent

    class Data:                                            # torch_geometric.data.data.Data: one GlobalStorage in _store
        def __init__(self, mapping):
            self.__dict__["_store"] = GlobalStorage(mapping, self)
    GlobalStorage.__module__, GlobalStorage.__qualname__ = "torch_geometric.data.storage", "GlobalStorage"
    Data.__module__, Data.__qualname__ = "torch_geometric.data.data", "Data"
    mods["torch_geometric.data.storage"].GlobalStorage = GlobalStorage
    mods["torch_geometric.data.data"].Data = Data
    graphs = [ds[i] for i in range(len(ds))]
    e = graphs[0].edge_index.shape[1]
    n = graphs[0].x.shape[0]
    mapping = {"x": torch.cat([g.x for g in graphs]), "ensemble": torch.cat([g.ensemble for g in graphs]),
               "y": torch.cat([g.y for g in graphs]), "edge_index": torch.cat([g.edge_index for g in graphs], dim=1),
               "edge_attr": torch.cat([g.edge_attr for g in graphs])}
    k = len(graphs)
    slices = {"x": torch.arange(k + 1) * n, "ensemble": torch.arange(k + 1) * n, "y": torch.arange(k + 1) * n,
              "edge_index": torch.arange(k + 1) * e, "edge_attr": torch.arange(k + 1) * e}
    path = tmp_path / "EUPPBench_24h_train_rf.pt"
    sys.modules.update(mods)
    try:
        torch.save((Data(mapping), slices), path)
    finally:
        for name in mods:
            sys.modules.pop(name, None)
    assert "torch_geometric" not in sys.modules
    got = EUPPBench(root_raw=str(tmp_path), root_processed=str(tmp_path), leadtime="24h", max_dist=250.0, split="train_rf")
    assert len(got) == k
    for i in range(k):
        for key in ("x", "ensemble", "y", "edge_index", "edge_attr"):
            a, b = getattr(got[i], key), getattr(graphs[i], key)
            assert torch.equal(torch.nan_to_num(a.reshape(b.shape), nan=-77.0), torch.nan_to_num(b, nan=-77.0)), (i, key)   # (y holds NaN)
    with pytest.raises(ValueError):
        EUPPBench(str(tmp_path), str(tmp_path), "24h", 100.0, split="nope")              # utils/dataset.py:52-53
    with pytest.raises(FileNotFoundError):
        EUPPBench(str(tmp_path), str(tmp_path), "72h", 100.0, split="train_rf")


@pytest.mark.parametrize("m,h,row_tile", [(976, 128, 8), (7808, 128, 32), (15616, 128, 64), (2928, 512, 64), (4880, 256, 64)])
def test_row_tile_does_not_identify_the_tensor_core_path(m, h, row_tile):
    """rc_gemm_row_tile / rc_gemm_tc_workspace are host-side answers (no device needed).  Below 16 384 rows the SIMT
    Linear also picks 64-row tiles once they fill the SMs, the statistics tile of the tensor-core kernels; only
    rc_gemm_tc_workspace > 0 says that the tensor-core kernels - the ones that write rc_gemm.a_out for every operand
    prologue - will run (kernels.gine_layer_fwd / _bwd choose their weight-gradient operands by it)."""
    from raincast_gnn_b200 import kernels as K
    assert K.gemm_row_tile(m, h, h) == row_tile
    assert not K.gemm_on_tensor_cores(m, h, h)
    if os.environ.get("RC_GEMM_TC", "1") != "0":
        assert K.gemm_on_tensor_cores(16384, h, h) and K.gemm_row_tile(16384, h, h) == 64
