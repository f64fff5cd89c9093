"""Block-level forward / backward schedules over the C ABI (no autograd here).

Every function takes plain CUDA tensors, allocates its outputs / saved activations with torch (the
library never allocates) and issues librc_b200.so calls on the current CUDA stream, so a whole
training step built from these blocks can be captured in one CUDA graph.  The autograd Functions in
functional.py and the graphed engine in engine.py are both thin callers of this file.

Reference arithmetic: models/gnn.py:10-141 (DeepSetEncoder, ResGnn, GNN), PyG GINEConv,
torch.nn.BatchNorm1d / Linear; see include/rc_b200.h for the per-kernel citations.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math

import torch

from . import _lib
from ._lib import (RC_A_RED, RC_A_ROW, RC_B_COL, RC_B_RED, RC_EPI_BN_RELU_BWD, RC_EPI_BN_STATS, RC_EPI_MASK_POS,
                   RC_EPI_ADD_RES, RC_EPI_NONE, RC_EPI_RELU, RC_EPI_RELU_RES, RC_OP_AFFINE2, RC_OP_BITMASK, RC_OP_BN_RELU,
                   RC_OP_GINE_AGGR, RC_OP_NONE)

import os

SAVE_OPERANDS = os.environ.get("RC_SAVE_OPERANDS", "1") != "0"   # large graphs: transformed GEMM operands written out once
FUSE_GINE = os.environ.get("RC_FUSE_GINE", "1") != "0"     # 0: separate aggregation kernel (A/B measurements, tests)
BN_EPS = 1e-5          # torch.nn.BatchNorm1d defaults (models/gnn.py:23)
BN_MOMENTUM = 0.1
_SM = 148


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _Side:
    """Optional second stream for work that is off the critical path of backward (weight / bias gradients and
    their reductions).  Enabled by the graphed engine: inside a CUDA-graph capture the event waits become graph
    edges, so the weight-gradient GEMMs run concurrently with the data-gradient chain on otherwise idle SMs."""
    stream = None
    alt = None             # a third stream for the tail of backward (DeepSets / dim_red weight gradients), see on_side(alt=True)
    alt_dirty = False      # work was issued on `alt` that the next reduction / join has to wait for
    keep: list = []
    pending = None         # the step's GradSink (grad_sink): partials a block left for the next block's flush


SIDE = _Side()


class _MaskDump:
    """Test instrumentation (include/rc_b200.h, "Test instrumentation"): while `active` is a dict, the backward blocks
    record the ReLU decisions they take - "phi" (bits [M*Em, H/32]), "rho" (bool [M, H]) and, per GINE layer in the order
    the backward visits them (last layer first), {"msg": bits per transpose slot, "bn": bits, "out": bits} under
    "layers" - so that tests/test_gpu_masked_parity.py can force them into the float64 oracle."""
    active = None


MASKS = _MaskDump()


def unpack_bits(bits, n):
    """bits int32 [rows, ceil(n/32)] -> bool [rows, n]."""
    words = bits.to(torch.int64) & 0xFFFFFFFF
    return ((words.unsqueeze(-1) >> torch.arange(32, device=bits.device)) & 1).reshape(bits.shape[0], -1)[:, :n].bool()


def _new(shape, dtype=torch.float32, device=None):
    return torch.empty(shape, dtype=dtype, device=device)


def _new_like(ref):
    return _new(ref.shape, ref.dtype, ref.device)


@contextlib.contextmanager
def on_side(*inputs, alt: bool = False):
    """Run the enclosed launches on the side stream, ordered after everything issued so far on the current stream.
    `inputs` (tensors produced on the main stream and read here) are kept alive until join_side().
    alt=True: on the second side stream when the engine provides one.  The weight-gradient GEMMs at the END of backward
    (dim_red, DeepSets) are the tail of the step - the optimiser waits for them - and independent of one another:
    alternating two streams overlaps their launch-to-finish latencies.  GradSink.flush / join_side wait for both."""
    if SIDE.stream is None:
        yield
        return
    s = SIDE.alt if (alt and SIDE.alt is not None) else SIDE.stream
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    s.wait_event(ev)
    SIDE.keep.extend(inputs)
    if s is SIDE.alt:
        SIDE.alt_dirty = True
    with torch.cuda.stream(s):
        yield


def _wait_alt():
    """Order the current stream behind the second side stream's work (before a reduction of partials it produced)."""
    if SIDE.alt is not None and SIDE.alt_dirty and torch.cuda.current_stream() != SIDE.alt:
        ev = torch.cuda.Event()
        ev.record(SIDE.alt)
        torch.cuda.current_stream().wait_event(ev)
        SIDE.alt_dirty = False


def join_side():
    if SIDE.stream is None:
        return
    if SIDE.pending is not None:
        SIDE.pending.flush()              # (every block sequence of the engine ends on a flushing block: normally empty)
        SIDE.pending = None
    ev = torch.cuda.Event()
    ev.record(SIDE.stream)
    torch.cuda.current_stream().wait_event(ev)
    _wait_alt()
    SIDE.keep.clear()


def operand(t, ld, op=RC_OP_NONE, p=(None, None, None, None), aux=None, ld_aux=0, bits=None, ld_bits=0, idx0=None, idx1=None):
    return _lib.rc_operand(_lib.ptr(t), ld, op, _lib.ptr(p[0]), _lib.ptr(p[1]), _lib.ptr(p[2]), _lib.ptr(p[3]),
                           _lib.ptr(aux), ld_aux, _lib.ptr(bits), ld_bits, _lib.ptr(idx0), _lib.ptr(idx1))


def gemm(m, n, k, a: _lib.rc_operand, b: _lib.rc_operand, d, ldd, *, a_layout=RC_A_ROW, b_layout=RC_B_COL, bias=None,
         bias_scale=1.0, epi=RC_EPI_NONE, res=None, ld_res=0, bits_out=None, ld_bits_out=0, e_aux=None, ld_e_aux=0,
         e_p=(None, None, None, None), stats=None, splits=1, split_stride=0, colsum_a=None, a2=None, lda2=0, b2=None,
         ldb2=0, k2=0, rows_per_warp=0, a_out=None, ld_a_out=0, b_static=None, run=True):
    if b_static is None:            # activation GEMMs: B is a parameter, fetched before the wait on the previous kernel
        b_static = a_layout == RC_A_ROW
    g = _lib.rc_gemm(m, n, k, a_layout, b_layout, a, b, _lib.ptr(a2), lda2, _lib.ptr(b2), ldb2, k2, _lib.ptr(d), ldd,
                     _lib.ptr(bias), bias_scale, epi, _lib.ptr(res), ld_res, _lib.ptr(bits_out), ld_bits_out,
                     _lib.ptr(e_aux), ld_e_aux, _lib.ptr(e_p[0]), _lib.ptr(e_p[1]), _lib.ptr(e_p[2]), _lib.ptr(e_p[3]),
                     _lib.ptr(stats), splits, split_stride, _lib.ptr(colsum_a), rows_per_warp, None, 0, _lib.ptr(a_out), ld_a_out, int(b_static))
    if run:
        L = _lib.lib()
        ws_bytes = int(L.rc_gemm_tc_workspace(C.byref(g)))     # > 0: large activation GEMM -> tensor cores (3xTF32)
        if ws_bytes:
            ws = _new(ws_bytes, torch.uint8, d.device)          # pre-split weight blocks, needed for this call only
            g.tc_ws, g.tc_ws_bytes = ws.data_ptr(), ws_bytes
        _lib.check(L.rc_gemm_run(C.byref(g), _stream(d)), "rc_gemm_run")
    return g


def gemm_on_tensor_cores(m, n, k) -> bool:
    """True when gemm() runs an m x n activation GEMM over k on the tcgen05 kernels (rc_gemm_tc_workspace > 0).  Only that
    path writes `a_out` for the BatchNorm / bit-mask / affine prologues; the SIMT kernels write it for RC_OP_GINE_AGGR only."""
    g = _lib.rc_gemm()
    g.m, g.n, g.k, g.splits = m, n, k, 1
    g.a_layout, g.b_layout = RC_A_ROW, RC_B_COL
    return bool(k) and int(_lib.lib().rc_gemm_tc_workspace(C.byref(g))) > 0


def gemm_row_tile(m, n, k=0) -> int:
    """Rows per statistics tile of the forward GEMM gemm() would launch for an m x n output over k."""
    g = _lib.rc_gemm()
    g.m, g.n, g.k, g.splits = m, n, k, 1
    g.a_layout, g.b_layout = RC_A_ROW, RC_B_COL
    if k and int(_lib.lib().rc_gemm_tc_workspace(C.byref(g))):
        return 64
    return int(_lib.lib().rc_gemm_row_tile(C.byref(g)))


# --------------------------------------------------------------------------------------------- Linear
def linear_fwd(x, w, b, *, relu=False, bias_scale=1.0):
    """y = [relu](x @ w^T + bias_scale*b);  x [M,K], w [N,K]."""
    m, k = x.shape
    n = w.shape[0]
    y = _new((m, n), torch.float32, x.device)
    gemm(m, n, k, operand(x, k), operand(w, k), y, n, bias=b, bias_scale=bias_scale,
         epi=RC_EPI_RELU if relu else RC_EPI_NONE)
    return y


def linear_bwd_data(dy, w, *, mask_pos=None, w_ld=None, w_col0=0, k=None):
    """dx = dy @ w[:, w_col0:w_col0+k]  (optionally masked by mask_pos > 0: ReLU backward).  dy [M,N], w [N,Kw]."""
    m, n = dy.shape
    w_ld = w.shape[1] if w_ld is None else w_ld
    k = w.shape[1] - w_col0 if k is None else k
    dx = _new((m, k), torch.float32, dy.device)
    wv = w if w_col0 == 0 else w.reshape(-1)[w_col0:]
    gemm(m, k, n, operand(dy, n), operand(wv, w_ld), dx, k, b_layout=RC_B_RED,
         epi=RC_EPI_MASK_POS if mask_pos is not None else RC_EPI_NONE, e_aux=mask_pos, ld_e_aux=k)
    return dx


def choose_splits(red_len: int, out_rows: int, out_cols: int) -> int:
    """Reduction splits for a weight-gradient GEMM: enough CTAs to fill the SMs, down to ONE 32-long reduction slice
    per CTA.  At the reference shape (976 rows) a slice costs a full global -> shared round trip (~3 us) whatever the
    tile does with it, so the 31 slices run side by side instead of 2 per CTA: the step went from 430 us to 389 us
    (the weight-gradient stream had become the critical path of backward)."""
    tiles = max(1, math.ceil(red_len / 32))
    out_tiles = math.ceil(out_rows / 64) * math.ceil(out_cols / 128)
    want = max(1, math.ceil(2 * _SM / out_tiles))
    return int(max(1, min(want, tiles)))


class GradSink:
    """Collects the split partials of one backward block and finishes them with one rc_reduce_segments launch."""

    def __init__(self, device):
        self.device = device
        self.segs = []
        self.keep = []

    def add(self, src, dst, stride, parts, n, scale=1.0, accumulate=False, row_len=0, dst_ld=0):
        self.segs.append(_lib.rc_reduce_seg(src.data_ptr(), dst.data_ptr(), stride, parts, n, scale, int(accumulate),
                                            row_len, dst_ld))
        self.keep.append((src, dst))

    def flush(self):
        if self.segs:
            _wait_alt()
            arr = (_lib.rc_reduce_seg * len(self.segs))(*self.segs)
            _lib.check(_lib.lib().rc_reduce_segments(arr, len(self.segs), torch.cuda.current_stream(self.device).cuda_stream),
                       "rc_reduce_segments")
        if SIDE.stream is not None:
            # partials written on one side stream and reduced on another: they must not go back to the allocator (and be
            # handed to a later launch on the stream that wrote them) before the step's streams have been joined
            SIDE.keep.extend(self.keep)
        self.segs, self.keep = [], []


def grad_sink(device):
    """The GradSink of a backward block.  Inside an engine step (side streams set) the blocks share ONE sink, so that a
    block may leave its partials to the next block's rc_reduce_segments launch (`flush=False`: the head's go with the last
    GINE layer's, dim_red's with the DeepSets block's) instead of paying a launch of its own."""
    if SIDE.stream is None:
        return GradSink(device)
    if SIDE.pending is None:
        SIDE.pending = GradSink(device)
    return SIDE.pending


def linear_bwd_weight(dy_op: _lib.rc_operand, x_op: _lib.rc_operand, m, n, k, dw, db, sink: GradSink, *, dw_ld=None,
                      bias_scale=1.0, x2=None, dw2=None):
    """dw[N, k] = dy^T @ x (+ db[N] = bias_scale * column sums of dy).  dy stored [M,N], x stored [M,k].
    `dw` may be a column block of a wider matrix (dw_ld = its row stride)."""
    dev = dw.device
    splits = int(_lib.lib().rc_gemm_tc_wgrad_splits(n, k, m)) or choose_splits(m, n, k)
    dw_ld = k if dw_ld is None else dw_ld
    direct = splits == 1 and dw_ld == k and bias_scale == 1.0
    if direct:
        gemm(n, k, m, dy_op, x_op, dw, k, a_layout=RC_A_RED, b_layout=RC_B_RED, colsum_a=db)
        return
    part = _new((splits, n, k), torch.float32, dev)
    cs = _new((splits, n), torch.float32, dev) if db is not None else None
    gemm(n, k, m, dy_op, x_op, part, k, a_layout=RC_A_RED, b_layout=RC_B_RED, splits=splits, split_stride=n * k,
         colsum_a=cs)
    if dw_ld == k:
        sink.add(part, dw, n * k, splits, n * k)
    else:                                           # column block of a wider weight matrix
        sink.add(part, dw, n * k, splits, n * k, row_len=k, dst_ld=dw_ld)
    if db is not None:
        sink.add(cs, db, n, splits, n, scale=bias_scale)


# --------------------------------------------------------------------------------------------- DeepSets
def deepsets_fwd(P, ens, bf16: bool = False):
    """rho(sum_e phi(ens[:, e])) with the second phi Linear hoisted behind the sum (models/gnn.py:64-68).
    bf16=True: the member contraction runs with bf16 operands / fp32 accumulation on the tensor cores (config 5)."""
    m, em, f = ens.shape
    h = P["phi0_w"].shape[0]
    L = _lib.lib()
    pooled = _new((m, h), torch.float32, ens.device)
    fn = L.rc_deepsets_pool_fwd_bf16 if bf16 else L.rc_deepsets_pool_fwd
    _lib.check(fn(ens.data_ptr(), P["phi0_w"].data_ptr(), P["phi0_b"].data_ptr(), pooled.data_ptr(),
                  m, em, f, h, _stream(ens)), "rc_deepsets_pool_fwd")
    s2 = linear_fwd(pooled, P["phi2_w"], P["phi2_b"], bias_scale=float(em))       # sum_e (h_e W^T + b)
    r1 = linear_fwd(s2, P["rho0_w"], P["rho0_b"], relu=True)
    emb = linear_fwd(r1, P["rho2_w"], P["rho2_b"])
    return emb, (ens, pooled, s2, r1, bf16)


def deepsets_bwd(P, saved, d_emb, G, mask_out=None):
    ens, pooled, s2, r1, bf16 = saved
    m, em, f = ens.shape
    h = P["phi0_w"].shape[0]
    L = _lib.lib()
    dev = ens.device
    sink = grad_sink(dev)
    ho = P["rho2_w"].shape[0]
    # rho[2]
    with on_side(d_emb):
        linear_bwd_weight(operand(d_emb, ho), operand(r1, h), m, ho, h, G["rho2_w"], G["rho2_b"], sink)
    d_r1 = linear_bwd_data(d_emb, P["rho2_w"], mask_pos=r1)
    # rho[0]
    with on_side(d_r1, alt=True):
        linear_bwd_weight(operand(d_r1, h), operand(s2, h), m, h, h, G["rho0_w"], G["rho0_b"], sink)
    d_s2 = linear_bwd_data(d_r1, P["rho0_w"])
    # phi[2] (after the pool): bias gradient carries the member count
    with on_side(d_s2):
        linear_bwd_weight(operand(d_s2, h), operand(pooled, h), m, h, h, G["phi2_w"], G["phi2_b"], sink, bias_scale=float(em))
    d_pooled = linear_bwd_data(d_s2, P["phi2_w"])
    # phi[0] + ReLU, per member
    nb = int(L.rc_deepsets_pool_bwd_nblocks(m, em, f, h))
    part = _new((nb, h * f + h), torch.float32, dev)
    if MASKS.active is not None:
        mask_out = torch.zeros((m * em, (h + 31) // 32), dtype=torch.int32, device=dev)
        MASKS.active["phi"] = mask_out
        MASKS.active["rho"] = r1 > 0
    _lib.check(L.rc_deepsets_pool_bwd(ens.data_ptr(), P["phi0_w"].data_ptr(), P["phi0_b"].data_ptr(), d_pooled.data_ptr(),
                                      part.data_ptr(), m, em, f, h, int(bf16), _lib.ptr(mask_out), _stream(ens)), "rc_deepsets_pool_bwd")
    sink.add(part, G["phi0_w"], h * f + h, nb, h * f)
    sink.add(part.reshape(-1)[h * f:], G["phi0_b"], h * f + h, nb, h)
    with on_side(part):
        sink.flush()


# --------------------------------------------------------------------------------------------- dim_red
def dimred_prepack(P, f, x=None):
    """Start of a step, on the side stream (overlaps the DeepSets kernels; the engine calls it):
    * dim_red.weight is [H, F + H] with F = 35: neither half of a row starts on a 16-byte boundary, and the GEMM falls
      back to scalar loads for the whole 128 x 163 operand.  Two strided copies split the weight into aligned halves
      wx [H, F padded to 4] and we [H, H]; the parameter itself keeps the reference layout.
    * with `x` (the batch's node features): xw = x @ wx^T + bias - the half of the Linear that does not need the
      DeepSets embedding - is computed here too, so that dimred_fwd on the critical path is one 128-long reduction
      slice that adds xw in its epilogue instead of two slices."""
    if SIDE.stream is None:
        return
    w = P["dimred_w"]
    n, ldw = w.shape
    pack = P.get("_dimred_pack")
    if pack is None:
        fpad = (f + 3) // 4 * 4
        pack = P["_dimred_pack"] = {"wx": torch.zeros((n, fpad), dtype=torch.float32, device=w.device),
                                    "we": torch.empty((n, ldw - f), dtype=torch.float32, device=w.device), "f": f}
    pack["xw"] = None
    with on_side(w):
        pack["wx"][:, :f].copy_(w.detach()[:, :f])
        pack["we"].copy_(w.detach()[:, f:])
        if x is not None:
            m = x.shape[0]
            xw = _new((m, n), torch.float32, x.device)
            gemm(m, n, f, operand(x, f), operand(pack["wx"], pack["wx"].shape[1]), xw, n, bias=P["dimred_b"],
                 b_static=False)     # wx was written by the copy kernel just before this one on the side stream
            pack["xw"] = xw
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
    pack["ready"] = ev


def _dimred_pack(P, f):
    """The aligned halves prepared for this step, or None (the eager path uses the parameter directly)."""
    pack = P.get("_dimred_pack")
    if pack is None or pack.get("ready") is None or pack["f"] != f:
        return None
    return pack


def dimred_fwd(P, x, emb):
    """Linear(cat([x, emb])) without materialising the cat: two reduction segments (models/gnn.py:134-135)."""
    m, f = x.shape
    h_in = emb.shape[1]
    w = P["dimred_w"]
    n, ldw = w.shape
    y = _new((m, n), torch.float32, x.device)
    pack = _dimred_pack(P, f)
    if pack is not None:
        torch.cuda.current_stream().wait_event(pack["ready"])
        if pack.get("xw") is not None:
            gemm(m, n, h_in, operand(emb, h_in), operand(pack["we"], h_in), y, n, epi=RC_EPI_ADD_RES, res=pack["xw"], ld_res=n)
        else:
            gemm(m, n, f, operand(x, f), operand(pack["wx"], pack["wx"].shape[1]), y, n, bias=P["dimred_b"],
                 a2=emb, lda2=h_in, b2=pack["we"], ldb2=h_in, k2=h_in)
    else:
        gemm(m, n, f, operand(x, f), operand(w, ldw), y, n, bias=P["dimred_b"],
             a2=emb, lda2=h_in, b2=w.reshape(-1)[f:], ldb2=ldw, k2=h_in)
    return y, (x, emb)


def dimred_bwd(P, saved, dy, G, *, flush: bool = True):
    x, emb = saved
    m, f = x.shape
    h_in = emb.shape[1]
    w = P["dimred_w"]
    n, ldw = w.shape
    sink = grad_sink(x.device)
    dw = G["dimred_w"]
    with on_side(dy, alt=True):
        linear_bwd_weight(operand(dy, n), operand(x, f), m, n, f, dw[:, :f], G["dimred_b"], sink, dw_ld=ldw)
    with on_side(dy):
        linear_bwd_weight(operand(dy, n), operand(emb, h_in), m, n, h_in, dw[:, f:], None, sink, dw_ld=ldw)
        if flush:
            sink.flush()
    pack = _dimred_pack(P, f)
    if pack is not None:
        pack["ready"] = None                 # one step only: the optimizer changes the weight next
        return linear_bwd_data(dy, pack["we"], w_ld=h_in, w_col0=0, k=h_in)
    return linear_bwd_data(dy, w, w_ld=ldw, w_col0=f, k=h_in)


# --------------------------------------------------------------------------------------------- GINE layer
def gine_aggr_fwd(x, graph, lin_w, lin_b, eps, out, *, tiled=None):
    """out = sum_{j -> i} relu(x_j + lin(e_ji)) + (1+eps) x_i  (PyG GINEConv message/aggregate, models/gnn.py:27-29).
    Large graphs go through the station tiles (shared-memory staging), small ones through the warp-per-row kernel;
    `tiled` forces the choice (tests)."""
    L = _lib.lib()
    m, h = x.shape
    tiles = graph.tiles(h) if tiled is None or tiled else None
    if tiled and tiles is None:
        raise _lib.RcError("tiled GINE aggregation requested but no station tiles apply to this graph / width")
    if tiles is not None:
        _lib.check(L.rc_gine_aggr_fwd_tiled(x.data_ptr(), C.byref(tiles[0].struct), lin_w.data_ptr(), lin_b.data_ptr(),
                                            eps.data_ptr(), out.data_ptr(), m, h, _stream(x)), "rc_gine_aggr_fwd_tiled")
    else:
        _lib.check(L.rc_gine_aggr_fwd(x.data_ptr(), graph.rowptr.data_ptr(), graph.col.data_ptr(), graph.attr.data_ptr(),
                                      lin_w.data_ptr(), lin_b.data_ptr(), eps.data_ptr(), out.data_ptr(), m, h, _stream(x)),
                   "rc_gine_aggr_fwd")
    return out


def gine_aggr_bwd(g, x, graph, lin_w, lin_b, eps, addend, dx, *, tiled=None):
    """dx and the per-block partials of d lin_w, d lin_b, d eps; returns (partials, nblocks)."""
    L = _lib.lib()
    m, h = x.shape
    tiles = graph.tiles(h) if tiled is None or tiled else None
    if tiled and tiles is None:
        raise _lib.RcError("tiled GINE aggregation requested but no station tiles apply to this graph / width")
    if tiles is not None:
        nb = int(L.rc_gine_aggr_bwd_tiled_nblocks(C.byref(tiles[1].struct), h))
        part = _new((nb, 3, h), torch.float32, x.device)
        _lib.check(L.rc_gine_aggr_bwd_tiled(g.data_ptr(), x.data_ptr(), C.byref(tiles[1].struct), lin_w.data_ptr(),
                                            lin_b.data_ptr(), eps.data_ptr(), _lib.ptr(addend), dx.data_ptr(), part.data_ptr(),
                                            m, h, _stream(x)), "rc_gine_aggr_bwd_tiled")
    else:
        nb = int(L.rc_gine_aggr_bwd_nblocks(m, h))
        part = _new((nb, 3, h), torch.float32, x.device)
        _lib.check(L.rc_gine_aggr_bwd(g.data_ptr(), x.data_ptr(), graph.t_rowptr.data_ptr(), graph.t_dst.data_ptr(),
                                      graph.t_attr.data_ptr(), lin_w.data_ptr(), lin_b.data_ptr(), eps.data_ptr(),
                                      _lib.ptr(addend), dx.data_ptr(), part.data_ptr(), m, h, _stream(x)), "rc_gine_aggr_bwd")
    return part, nb


def gine_layer_fwd(P, x, graph, *, first: bool, training: bool):
    """One ResGnn layer (models/gnn.py:39-44): y = relu(conv(x)) for the first layer, x + relu(conv(x)) after;
    conv = GINEConv(nn = Linear - BatchNorm1d - ReLU - Linear)."""
    L = _lib.lib()
    m, h = x.shape
    dev = x.device
    st = _stream(x)
    agg = _new_like(x)
    # small graphs (every reference-shape batch): the aggregation is the operand prologue of the first Linear - one kernel
    # for message + aggregation + (1+eps) x + Linear1 + BatchNorm tile statistics; large graphs aggregate over station
    # tiles (shared-memory staging) and run the Linear on the tensor cores
    fused = FUSE_GINE and m < 16384 and h % 4 == 0 and graph.tiles(h) is None
    a_op = (operand(x, h, RC_OP_GINE_AGGR, (P["lin_w"], P["lin_b"], P["eps"], None), aux=graph.attr, idx0=graph.rowptr, idx1=graph.col)
            if fused else operand(agg, h))
    a_out = agg if fused else None
    if not fused:
        gine_aggr_fwd(x, graph, P["lin_w"], P["lin_b"], P["eps"], agg)
    hid = P["nn0_w"].shape[0]
    t = _new((m, hid), torch.float32, dev)
    mean = _new(hid, torch.float32, dev)
    rstd = _new(hid, torch.float32, dev)
    if training:
        if m < 2:
            raise ValueError("Expected more than 1 value per channel when training (BatchNorm1d)")
        row_tile = gemm_row_tile(m, hid, h)
        tiles = math.ceil(m / row_tile)
        stats = _new((tiles, 2, hid), torch.float32, dev)
        gemm(m, hid, h, a_op, operand(P["nn0_w"], h), t, hid, bias=P["nn0_b"], epi=RC_EPI_BN_STATS, stats=stats, a_out=a_out, ld_a_out=h)
        _lib.check(L.rc_bn_stats_finalize(stats.data_ptr(), tiles, row_tile, m, hid, BN_EPS, BN_MOMENTUM, mean.data_ptr(),
                                          rstd.data_ptr(), P["bn_rm"].data_ptr(), P["bn_rv"].data_ptr(),
                                          P["bn_nbt"].data_ptr(), st), "rc_bn_stats_finalize")
    else:
        gemm(m, hid, h, a_op, operand(P["nn0_w"], h), t, hid, bias=P["nn0_b"], a_out=a_out, ld_a_out=h)
        _lib.check(L.rc_bn_eval_prepare(P["bn_rm"].data_ptr(), P["bn_rv"].data_ptr(), hid, BN_EPS, mean.data_ptr(),
                                        rstd.data_ptr(), st), "rc_bn_eval_prepare")
    out_dim = P["nn3_w"].shape[0]
    y = _new((m, out_dim), torch.float32, dev)
    words = math.ceil(out_dim / 32)
    bits = _new((m, words), torch.int32, dev)
    # large graphs (tensor-core Linear layers): u = relu(BN(t)) is written out by the GEMM's operand producer and the
    # backward's weight-gradient GEMM reads it back instead of recomputing it element by element
    # (asked of the library, not inferred from the row tile: a SIMT GEMM with 64-row tiles - 9.4k to 16k rows at H=128,
    # from 2.3k rows at H=512 - does not write u)
    u = _new((m, hid), torch.float32, dev) if (training and SAVE_OPERANDS and gemm_on_tensor_cores(m, out_dim, hid)) else None
    gemm(m, out_dim, hid, operand(t, hid, RC_OP_BN_RELU, (mean, rstd, P["bn_w"], P["bn_b"])), operand(P["nn3_w"], hid),
         y, out_dim, bias=P["nn3_b"], epi=RC_EPI_RELU if first else RC_EPI_RELU_RES, res=None if first else x,
         ld_res=h, bits_out=bits, ld_bits_out=words, a_out=u, ld_a_out=hid)
    return y, (x, agg, t, mean, rstd, bits, u)


def gine_layer_bwd(P, saved, graph, dy, G, *, first: bool, training: bool = True, need_dx: bool = True):
    x, agg, t, mean, rstd, bits, u = saved
    L = _lib.lib()
    m, h = x.shape
    hid = P["nn0_w"].shape[0]
    out_dim = P["nn3_w"].shape[0]
    words = bits.shape[1]
    dev = x.device
    st = _stream(x)
    sink = grad_sink(dev)
    do_op = operand(dy, out_dim, RC_OP_BITMASK, bits=bits, ld_bits=words)           # d o = dy * 1[o > 0]
    row_tile = gemm_row_tile(m, hid, out_dim)
    # tensor-core path: the activation GEMMs write their transformed operand out
    big = u is not None and gemm_on_tensor_cores(m, hid, out_dim) and gemm_on_tensor_cores(m, h, hid)
    # Linear2: d W2 = d o^T u,  u = relu(BN(t)) recomputed in the prologue (small graphs) or saved by the forward (large)
    if not big:
        with on_side(dy):
            linear_bwd_weight(do_op, operand(t, hid, RC_OP_BN_RELU, (mean, rstd, P["bn_w"], P["bn_b"])), m, out_dim, hid,
                              G["nn3_w"], G["nn3_b"], sink)
    # d z = (d o @ W2) * 1[BN(t) > 0], with the two BatchNorm column reductions in the epilogue
    tiles = math.ceil(m / row_tile)
    stats = _new((tiles, 2, hid), torch.float32, dev)
    dz = _new((m, hid), torch.float32, dev)
    do_buf = _new((m, out_dim), torch.float32, dev) if big else None
    gemm(m, hid, out_dim, do_op, operand(P["nn3_w"], hid), dz, hid, b_layout=RC_B_RED, epi=RC_EPI_BN_RELU_BWD, e_aux=t,
         ld_e_aux=hid, e_p=(mean, rstd, P["bn_w"], P["bn_b"]), stats=stats, a_out=do_buf, ld_a_out=out_dim)
    if big:
        with on_side(do_buf, u):
            linear_bwd_weight(operand(do_buf, out_dim), operand(u, hid), m, out_dim, hid, G["nn3_w"], G["nn3_b"], sink)
    c0 = _new(hid, torch.float32, dev)
    c1 = _new_like(c0)
    c2 = _new_like(c0)
    _lib.check(L.rc_bn_bwd_finalize(stats.data_ptr(), tiles, m, hid, int(training), P["bn_w"].data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    G["bn_w"].data_ptr(), G["bn_b"].data_ptr(), c0.data_ptr(), c1.data_ptr(), c2.data_ptr(), st),
               "rc_bn_bwd_finalize")
    dt_op = operand(dz, hid, RC_OP_AFFINE2, (c0, c1, c2, mean), aux=t, ld_aux=hid)  # d t = c0*dz + c1*(t-mean) + c2
    d_agg = _new((m, h), torch.float32, dev)
    if big:
        dt_buf = _new((m, hid), torch.float32, dev)
        gemm(m, h, hid, dt_op, operand(P["nn0_w"], h), d_agg, h, b_layout=RC_B_RED, a_out=dt_buf, ld_a_out=hid)
        with on_side(dt_buf):
            linear_bwd_weight(operand(dt_buf, hid), operand(agg, h), m, hid, h, G["nn0_w"], G["nn0_b"], sink)
    else:
        with on_side(dz, c0, c1, c2):
            linear_bwd_weight(dt_op, operand(agg, h), m, hid, h, G["nn0_w"], G["nn0_b"], sink)
        gemm(m, h, hid, dt_op, operand(P["nn0_w"], h), d_agg, h, b_layout=RC_B_RED)
    # aggregation backward (+ residual branch of layers > 0)
    dx = _new((m, h), torch.float32, dev)
    if MASKS.active is not None:
        msg_bits = torch.zeros((graph.num_edges, (h + 31) // 32), dtype=torch.int32, device=dev)
        bn_bits = torch.zeros((m, (hid + 31) // 32), dtype=torch.int32, device=dev)
        _lib.check(L.rc_debug_gine_msg_mask(x.data_ptr(), graph.t_rowptr.data_ptr(), graph.t_attr.data_ptr(), P["lin_w"].data_ptr(),
                                            P["lin_b"].data_ptr(), m, h, int(graph.tiles(h) is not None), msg_bits.data_ptr(), st),
                   "rc_debug_gine_msg_mask")
        _lib.check(L.rc_debug_bn_relu_mask(t.data_ptr(), hid, mean.data_ptr(), rstd.data_ptr(), P["bn_w"].data_ptr(),
                                           P["bn_b"].data_ptr(), m, hid, bn_bits.data_ptr(), st), "rc_debug_bn_relu_mask")
        MASKS.active.setdefault("layers", []).append({"msg": msg_bits, "bn": bn_bits, "out": bits})
    part, nb = gine_aggr_bwd(d_agg, x, graph, P["lin_w"], P["lin_b"], P["eps"], None if first else dy, dx)
    # the per-CTA partials [nb][d w_edge (h) | d b_edge (h) | d eps] join the layer's other split gradients: ONE
    # rc_reduce_segments launch per layer (float64, fixed order - the arithmetic of rc_gine_aggr_bwd_finalize)
    flat = part.reshape(-1)
    sink.add(part, G["lin_w"], 3 * h, nb, h)
    sink.add(flat[h:], G["lin_b"], 3 * h, nb, h)
    sink.add(flat[2 * h:], G["eps"], 3 * h, nb, 1)
    with on_side(part, d_agg):
        sink.flush()
    return dx if need_dx else None


# --------------------------------------------------------------------------------------------- head
def head_fwd(P, x):
    return linear_fwd(x, P["aggr_w"], P["aggr_b"]), (x,)


def head_bwd(P, saved, d_raw, G, *, flush: bool = True):
    (x,) = saved
    m, h = x.shape
    c = P["aggr_w"].shape[0]
    sink = grad_sink(x.device)
    with on_side(d_raw):
        linear_bwd_weight(operand(d_raw, c), operand(x, h), m, c, h, G["aggr_w"], G["aggr_b"], sink)
        if flush:
            sink.flush()
    return linear_bwd_data(d_raw, P["aggr_w"])


def head_crps_blocks(m: int, hidden: int) -> int:
    """CTAs of the fused head + CRPS kernel, 0 when it does not apply (large batches, other widths)."""
    return int(_lib.lib().rc_head_crps_blocks(int(m), int(hidden)))


def head_crps_fwd_bwd(P, x, y, kind, G, *, u=0.0, xi=0.5, t=5.0, loss_out=None, flush: bool = True):
    """Head Linear + links + CRPS + backward in ONE launch (small batches): returns (loss float64[1], d_x, n_valid).
    The weight / bias gradient partials are reduced on the side stream."""
    m, h = x.shape
    c = P["aggr_w"].shape[0]
    dev = x.device
    nb = head_crps_blocks(m, h)
    d_x = _new((m, h), torch.float32, dev)
    part = _new((nb, c * h + c), torch.float32, dev)
    lp = _new((nb,), torch.float64, dev)
    loss = loss_out if loss_out is not None else _new((1,), torch.float64, dev)
    n_valid = _new((1,), torch.int32, dev)
    _lib.check(_lib.lib().rc_head_crps_fwd_bwd(x.data_ptr(), P["aggr_w"].data_ptr(), P["aggr_b"].data_ptr(), y.data_ptr(), d_x.data_ptr(),
                                               part.data_ptr(), lp.data_ptr(), loss.data_ptr(), n_valid.data_ptr(), m, h, kind,
                                               float(u), float(xi), float(t), _stream(x)), "rc_head_crps_fwd_bwd")
    sink = grad_sink(dev)
    sink.add(part, G["aggr_w"], c * h + c, nb, c * h)
    sink.add(part.reshape(-1)[c * h:], G["aggr_b"], c * h + c, nb, c)
    if flush:
        with on_side(part, lp):
            sink.flush()
    else:
        SIDE.keep.extend((part, lp))
    return loss, d_x, n_valid


# --------------------------------------------------------------------------------------------- links + CRPS
def postprocess_fwd(raw, kind):
    post = _new_like(raw)
    _lib.check(_lib.lib().rc_postprocess_fwd(raw.data_ptr(), post.data_ptr(), raw.shape[0], kind, _stream(raw)), "rc_postprocess_fwd")
    return post


def postprocess_bwd(raw, d_post, kind):
    d_raw = _new_like(raw)
    _lib.check(_lib.lib().rc_postprocess_bwd(raw.data_ptr(), d_post.data_ptr(), d_raw.data_ptr(), raw.shape[0], kind, _stream(raw)),
               "rc_postprocess_bwd")
    return d_raw


def crps_fwd_bwd(pred, y, kind, *, raw_input=False, u=0.0, xi=0.5, t=5.0, need_grad=True, loss_out=None):
    """(loss float64 [1], d_pred or None, n_valid int32 [1]); d_pred already carries the 1/n_valid of the mean."""
    L = _lib.lib()
    m = pred.shape[0]
    dev = pred.device
    ws_bytes = int(L.rc_crps_workspace(m))
    ws = _new(ws_bytes, torch.uint8, dev)
    loss = loss_out if loss_out is not None else _new(1, torch.float64, dev)
    n_valid = _new(1, torch.int32, dev)
    d_pred = _new_like(pred) if need_grad else None
    _lib.check(L.rc_crps_fwd_bwd(pred.data_ptr(), y.data_ptr(), _lib.ptr(d_pred), loss.data_ptr(), n_valid.data_ptr(), m, kind,
                                 int(raw_input), float(u), float(xi), float(t), ws.data_ptr(), ws_bytes, _stream(pred)),
               "rc_crps_fwd_bwd")
    return loss, d_pred, n_valid
