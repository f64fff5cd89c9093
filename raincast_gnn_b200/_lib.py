"""ctypes binding of the C ABI declared in include/rc_b200.h (librc_b200.so, built in-tree by
raincast_gnn_b200/csrc/build.py).  There is NO fallback: if the library is missing or a call
fails, an exception is raised — the product never computes on the CPU or through PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RC_B200_LIB") or os.path.join(_HERE, "csrc", "librc_b200.so")   # RC_B200_LIB: instrumented builds (tools/)

RC_A_ROW, RC_A_RED = 0, 1
RC_B_COL, RC_B_RED = 0, 1
RC_OP_NONE, RC_OP_BN_RELU, RC_OP_BITMASK, RC_OP_AFFINE2, RC_OP_GINE_AGGR = 0, 1, 2, 3, 4
RC_EPI_NONE, RC_EPI_RELU, RC_EPI_RELU_RES, RC_EPI_BN_STATS, RC_EPI_MASK_POS, RC_EPI_BN_RELU_BWD, RC_EPI_ADD_RES = 0, 1, 2, 3, 4, 5, 6
RC_LOSS_NORMAL, RC_LOSS_MIXED_NORMAL, RC_LOSS_MIXED, RC_LOSS_MIXED_U = 0, 1, 2, 3

_fp = C.c_void_p


class RcError(RuntimeError):
    pass


class rc_csr(C.Structure):
    _fields_ = [(n, _fp) for n in ("rowptr", "col", "attr", "perm", "t_rowptr", "t_dst", "t_attr", "t_perm", "t_slot", "rev")]


class rc_gine_tiles(C.Structure):
    _fields_ = [("n_tiles", C.c_int32), ("max_staged", C.c_int32), ("max_block_bytes", C.c_int32), ("row_bytes", C.c_int32)] + [
        (n, _fp) for n in ("tile_stage_ptr", "tile_blk_ptr", "stage_id", "blocks", "sched")]


class rc_operand(C.Structure):
    _fields_ = [("ptr", _fp), ("ld", C.c_int), ("op", C.c_int), ("p0", _fp), ("p1", _fp), ("p2", _fp), ("p3", _fp),
                ("aux", _fp), ("ld_aux", C.c_int), ("bits", _fp), ("ld_bits", C.c_int), ("idx0", _fp), ("idx1", _fp)]


class rc_gemm(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("k", C.c_int), ("a_layout", C.c_int), ("b_layout", C.c_int),
                ("a", rc_operand), ("b", rc_operand),
                ("a2", _fp), ("lda2", C.c_int), ("b2", _fp), ("ldb2", C.c_int), ("k2", C.c_int),
                ("d", _fp), ("ldd", C.c_int), ("bias", _fp), ("bias_scale", C.c_float), ("epi", C.c_int),
                ("res", _fp), ("ld_res", C.c_int), ("bits_out", _fp), ("ld_bits_out", C.c_int),
                ("e_aux", _fp), ("ld_e_aux", C.c_int), ("e_p0", _fp), ("e_p1", _fp), ("e_p2", _fp), ("e_p3", _fp),
                ("stats", _fp), ("splits", C.c_int), ("split_stride", C.c_longlong), ("colsum_a", _fp),
                ("rows_per_warp", C.c_int), ("tc_ws", _fp), ("tc_ws_bytes", C.c_size_t), ("a_out", _fp), ("ld_a_out", C.c_int), ("b_static", C.c_int)]


class rc_reduce_seg(C.Structure):
    _fields_ = [("src", _fp), ("dst", _fp), ("stride", C.c_longlong), ("parts", C.c_int), ("n", C.c_int),
                ("scale", C.c_float), ("accumulate", C.c_int), ("row_len", C.c_int), ("dst_ld", C.c_int)]


_lib = None


def _declare(lib):
    i, f, p, ll, sz = C.c_int, C.c_float, _fp, C.c_longlong, C.c_size_t
    sig = {
        "rc_version": (i, []),
        "rc_last_error": (C.c_char_p, []),
        "rc_launch_count": (C.c_uint64, []),
        "rc_radius_graph_count_host": (i, [p, i, f, p]),
        "rc_radius_graph_fill_host": (i, [p, i, f, ll, p, p]),
        "rc_radius_graph_coords_count_host": (i, [p, i, C.c_double, p]),
        "rc_radius_graph_coords_fill_host": (i, [p, i, C.c_double, ll, p, p]),
        "rc_csr_build_host": (i, [p, p, ll, i, C.POINTER(rc_csr)]),
        "rc_csr_build_workspace": (sz, [ll, i]),
        "rc_csr_build": (i, [p, p, ll, i, C.POINTER(rc_csr), p, sz, p, p]),
        "rc_gine_aggr_fwd": (i, [p, p, p, p, p, p, p, p, i, i, p]),
        "rc_gine_aggr_bwd_nblocks": (i, [i, i]),
        "rc_gine_aggr_bwd": (i, [p, p, p, p, p, p, p, p, p, p, p, i, i, p]),
        "rc_gine_aggr_bwd_finalize": (i, [p, i, i, p, p, p, p]),
        "rc_gine_tiles_limits": (i, [i, p, p]),
        "rc_gine_tiles_build_host": (i, [p, p, p, i, ll, i, i, i, p, p, p, p, p, p, p, p, p, p]),
        "rc_gine_tiles_verify_host": (i, [p, p, p, i, ll, i, i, i, i, p, p, p, p]),
        "rc_gine_aggr_fwd_tiled": (i, [p, C.POINTER(rc_gine_tiles), p, p, p, p, i, i, p]),
        "rc_gine_aggr_bwd_tiled_nblocks": (i, [C.POINTER(rc_gine_tiles), i]),
        "rc_gine_aggr_bwd_tiled": (i, [p, p, C.POINTER(rc_gine_tiles), p, p, p, p, p, p, i, i, p]),
        "rc_gemm_row_tile": (i, [C.POINTER(rc_gemm)]),
        "rc_gemm_run": (i, [C.POINTER(rc_gemm), p]),
        "rc_gemm_tc_workspace": (sz, [C.POINTER(rc_gemm)]),
        "rc_gemm_tc_wgrad_splits": (i, [i, i, i]),
        "rc_bn_stats_finalize": (i, [p, i, i, i, i, f, f, p, p, p, p, p, p]),
        "rc_bn_eval_prepare": (i, [p, p, i, f, p, p, p]),
        "rc_bn_bwd_finalize": (i, [p, i, i, i, i, p, p, p, p, p, p, p, p, p]),
        "rc_reduce_segments": (i, [C.POINTER(rc_reduce_seg), i, p]),
        "rc_deepsets_pool_fwd": (i, [p, p, p, p, i, i, i, i, p]),
        "rc_deepsets_pool_fwd_bf16": (i, [p, p, p, p, i, i, i, i, p]),
        "rc_deepsets_pool_bwd_nblocks": (i, [i, i, i, i]),
        "rc_deepsets_pool_bwd": (i, [p, p, p, p, p, i, i, i, i, i, p, p]),
        "rc_postprocess_fwd": (i, [p, p, i, i, p]),
        "rc_postprocess_bwd": (i, [p, p, p, i, i, p]),
        "rc_crps_workspace": (sz, [i]),
        "rc_crps_fwd_bwd": (i, [p, p, p, p, p, i, i, i, f, f, f, p, sz, p]),
        "rc_head_crps_blocks": (i, [i, i]),
        "rc_head_crps_fwd_bwd": (i, [p, p, p, p, p, p, p, p, p, i, i, i, f, f, f, p]),
        "rc_adamw_step": (i, [p, p, p, p, p, ll, f, f, f, f, f, f, p]),
        "rc_gather_dates": (i, [p, p, p, p, i, i, ll, ll, ll, p, p, p, p, p]),
        "rc_gather_dates_step": (i, [p, p, p, p, i, p, p, i, i, ll, ll, ll, p, p, p, p, p]),
        "rc_p2p_barrier": (i, [p, p, i, i, i, p, p]),
        "rc_p2p_adamw_step": (i, [p, p, i, p, p, p, ll, f, f, f, f, f, p]),
        "rc_p2p_step": (i, [p, p, p, p, i, i, p, p, p, ll, f, f, f, f, f, p, p]),
        "rc_p2p_wait_done": (i, [p, p, i, i, p, p]),
        "rc_p2p_flag_scope": (i, [i]),
        "rc_debug_gine_msg_mask": (i, [p, p, p, p, p, i, i, i, p, p]),
        "rc_debug_bn_relu_mask": (i, [p, i, p, p, p, p, i, i, p, p]),
        "rc_debug_fma_peak": (i, [p, i, p, p]),
        "rc_debug_tc_trace": (None, [p]),
        "rc_debug_p2p_trace": (None, [p]),
        "rc_debug_ds_trace": (None, [p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return sig


EXPORTS = None


def lib():
    """The loaded library; raises RcError if it has not been built (python -m raincast_gnn_b200.csrc.build)."""
    global _lib, EXPORTS
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RcError(f"{LIB_PATH} is missing: build it with `python -m raincast_gnn_b200.csrc.build` "
                          "(there is no CPU / PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        EXPORTS = sorted(_declare(handle).keys())
        _lib = handle
    return _lib


def check(code: int, what: str = ""):
    if code != 0:
        msg = lib().rc_last_error().decode(errors="replace")
        raise RcError(f"{what or 'rc call'} failed with code {code}: {msg}")


def launch_count() -> int:
    return int(lib().rc_launch_count())


def ptr(t):
    """Device/host address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RcError("raincast_gnn_b200 kernels need CUDA tensors; there is no CPU path "
                          "(move the model and the batch to a cuda device)")


def f32c(t: torch.Tensor) -> torch.Tensor:
    """float32, contiguous, 16-byte aligned view/copy of t (the kernels use 128-bit loads)."""
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()
