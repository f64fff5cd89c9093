"""Time the Linear-layer GEMMs of configs 4 / 5 (100k rows) on the tensor-core path; run with RC_GEMM_TC=0 for the SIMT kernels.
    python tools/time_gemm_tc.py [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
m = 100000
h = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(m, h, generator=g, device=dev)
dy = torch.randn(m, h, generator=g, device=dev)
w = torch.randn(h, h, generator=g, device=dev) / h ** 0.5
b = torch.randn(h, generator=g, device=dev)
vec = [torch.rand(h, generator=g, device=dev) + 0.5 for _ in range(4)]
y = torch.empty(m, h, device=dev)
bits = torch.zeros(m, h // 32, dtype=torch.int32, device=dev)
stats = torch.empty((m + 127) // 128 + 8000, 2, h, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, flops, bytes_):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    print(f"{name:44s} {t:8.1f} us   {flops / t * 1e-6:7.1f} TFLOP/s   {bytes_ / t * 1e-3:7.1f} GB/s")


fl = 2.0 * m * h * h
by = 2.0 * m * h * 4
timeit("forward bias+relu", lambda: K.gemm(m, h, h, K.operand(x, h), K.operand(w, h), y, h, bias=b, epi=K.RC_EPI_RELU), fl, by)
timeit("forward BN stats", lambda: K.gemm(m, h, h, K.operand(x, h), K.operand(w, h), y, h, bias=b, epi=K.RC_EPI_BN_STATS, stats=stats), fl, by)
timeit("forward BN+ReLU prologue, relu+res+bits", lambda: K.gemm(m, h, h, K.operand(x, h, K.RC_OP_BN_RELU, vec), K.operand(w, h), y, h, bias=b,
                                                                  epi=K.RC_EPI_RELU_RES, res=dy, ld_res=h, bits_out=bits, ld_bits_out=h // 32), fl, by * 1.5)
do_op = K.operand(dy, h, K.RC_OP_BITMASK, bits=bits, ld_bits=h // 32)
timeit("backward-data bitmask -> BN+ReLU bwd", lambda: K.gemm(m, h, h, do_op, K.operand(w, h), y, h, b_layout=K.RC_B_RED, epi=K.RC_EPI_BN_RELU_BWD,
                                                               e_aux=x, ld_e_aux=h, e_p=vec, stats=stats), fl, by * 1.5)
dt_op = K.operand(dy, h, K.RC_OP_AFFINE2, vec, aux=x, ld_aux=h)
timeit("backward-data affine2 prologue", lambda: K.gemm(m, h, h, dt_op, K.operand(w, h), y, h, b_layout=K.RC_B_RED), fl, by * 1.5)
dw, db = torch.empty(h, h, device=dev), torch.empty(h, device=dev)


def wgrad(a_op, b_op):
    sink = K.GradSink(dev)
    K.linear_bwd_weight(a_op, b_op, m, h, h, dw, db, sink)
    sink.flush()


timeit("weight grad plain (+reduce)", lambda: wgrad(K.operand(dy, h), K.operand(x, h)), fl, by)
timeit("weight grad affine2 x plain (+reduce)", lambda: wgrad(dt_op, K.operand(x, h)), fl, by * 1.5)
timeit("weight grad bitmask x BN+ReLU (+reduce)", lambda: wgrad(do_op, K.operand(x, h, K.RC_OP_BN_RELU, vec)), fl, by)
