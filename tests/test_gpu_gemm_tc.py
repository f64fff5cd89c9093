"""GPU parity of the tensor-core (tcgen05, 3xTF32) path of rc_gemm_run - taken for activation GEMMs with >= 16384 rows and
weight-gradient GEMMs over >= 16384 samples (BASELINE.json configs 4 / 5) - against float64 torch restatements of the
same fused op.  Tolerance: max|a-b| / max|b| <= 1e-5 (fp32 north_star budget: the 3xTF32 split must hold it)."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _np(t):
    return t.detach().double().cpu().numpy()


def _tc_used(fn):
    """Run fn and return the kernels' launch-count delta (a tensor-core rows GEMM is one launch with the weight tile in
    TMEM, pack + kernel = 2 otherwise; the SIMT path is also 1, so the callers also check the workspace query)."""
    from raincast_gnn_b200 import _lib
    before = _lib.launch_count()
    fn()
    torch.cuda.synchronize()
    return _lib.launch_count() - before


@pytest.mark.parametrize("m,n,k", [(20000, 128, 128), (16384, 128, 35), (33333, 5, 128), (17000, 256, 256), (16500, 512, 512), (20001, 128, 163)])
def test_tc_forward_bias_relu(dev, m, n, k):
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator(device=dev).manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g, device=dev)
    w = torch.randn(n, k, generator=g, device=dev)
    b = torch.randn(n, generator=g, device=dev)
    want = torch.relu(x.double() @ w.double().T + 2.0 * b.double())
    y = torch.full((m, n), float("nan"), device=dev)
    launches = _tc_used(lambda: K.gemm(m, n, k, K.operand(x, k), K.operand(w, k), y, n, bias=b, bias_scale=2.0, epi=K.RC_EPI_RELU))
    assert launches in (1, 2)
    assert K.gemm_row_tile(m, n, k) == 64, "the tensor-core path was not taken"
    assert rel_err(_np(y), _np(want)) < TOL


@pytest.mark.parametrize("m,n,k", [(20000, 128, 128), (16390, 128, 5), (18000, 35, 128), (16500, 512, 512)])
def test_tc_backward_data_mask(dev, m, n, k):
    """dx[m, n] = dy[m, k] @ w[k, n]  (w stored [k][n]: RC_B_RED), masked by aux > 0."""
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator(device=dev).manual_seed(7 + m + n + k)
    dy = torch.randn(m, k, generator=g, device=dev)
    w = torch.randn(k, n, generator=g, device=dev)
    aux = torch.randn(m, n, generator=g, device=dev)
    want = (dy.double() @ w.double()) * (aux > 0)
    out = torch.full((m, n), float("nan"), device=dev)
    if k < 32:
        pytest.skip("reductions shorter than one 32-wide block stay on the SIMT kernel")
    launches = _tc_used(lambda: K.gemm(m, n, k, K.operand(dy, k), K.operand(w, n), out, n, b_layout=K.RC_B_RED,
                                       epi=K.RC_EPI_MASK_POS, e_aux=aux, ld_e_aux=n))
    assert launches in (1, 2) and K.gemm_row_tile(m, n, k) == 64
    assert rel_err(_np(out), _np(want)) < TOL


def test_tc_two_segments_unaligned(dev):
    """Linear(cat([x, emb])) as two reduction segments, x [M, 35] not 16-byte aligned per row (models/gnn.py:134-135)."""
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator(device=dev).manual_seed(5)
    m, f, h = 20011, 35, 128
    x = torch.randn(m, f, generator=g, device=dev)
    emb = torch.randn(m, h, generator=g, device=dev)
    w = torch.randn(h, f + h, generator=g, device=dev) * 0.1
    b = torch.randn(h, generator=g, device=dev)
    y, _ = K.dimred_fwd({"dimred_w": w, "dimred_b": b}, x, emb)
    want = torch.cat([x, emb], 1).double() @ w.double().T + b.double()
    assert rel_err(_np(y), _np(want)) < TOL


@pytest.mark.parametrize("m,h", [(20000, 128), (16385, 256)])
def test_tc_node_mlp_forward_chain(dev, m, h):
    """Linear -> BatchNorm1d(train) statistics -> ReLU -> Linear (+ReLU, residual, bit mask), all on the tensor-core path."""
    from raincast_gnn_b200 import _lib, kernels as K
    g = torch.Generator(device=dev).manual_seed(m + h)
    x = torch.randn(m, h, generator=g, device=dev) + 3.0
    w1 = torch.randn(h, h, generator=g, device=dev) / math.sqrt(h)
    b1 = torch.randn(h, generator=g, device=dev)
    w2 = torch.randn(h, h, generator=g, device=dev) / math.sqrt(h)
    b2 = torch.randn(h, generator=g, device=dev)
    gam = torch.rand(h, generator=g, device=dev) + 0.5
    bet = torch.randn(h, generator=g, device=dev)
    res = torch.randn(m, h, generator=g, device=dev)
    t_ref = x.double() @ w1.double().T + b1.double()
    mean_ref, var_ref = t_ref.mean(0), t_ref.var(0, unbiased=False)
    u_ref = torch.relu((t_ref - mean_ref) / torch.sqrt(var_ref + 1e-5) * gam.double() + bet.double())
    o_ref = u_ref @ w2.double().T + b2.double()
    y_ref = res.double() + torch.relu(o_ref)
    row_tile = K.gemm_row_tile(m, h, h)
    assert row_tile == 64
    tiles = math.ceil(m / row_tile)
    t = torch.empty(m, h, device=dev)
    stats = torch.empty(tiles, 2, h, device=dev)
    K.gemm(m, h, h, K.operand(x, h), K.operand(w1, h), t, h, bias=b1, epi=K.RC_EPI_BN_STATS, stats=stats)
    mean, rstd = torch.empty(h, device=dev), torch.empty(h, device=dev)
    rm, rv, nbt = torch.zeros(h, device=dev), torch.ones(h, device=dev), torch.zeros((), dtype=torch.long, device=dev)
    _lib.check(_lib.lib().rc_bn_stats_finalize(stats.data_ptr(), tiles, row_tile, m, h, 1e-5, 0.1, mean.data_ptr(), rstd.data_ptr(),
                                               rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert rel_err(_np(t), _np(t_ref)) < TOL
    assert rel_err(_np(mean), _np(mean_ref)) < TOL
    assert rel_err(_np(rstd), _np(1 / torch.sqrt(var_ref + 1e-5))) < TOL
    words = math.ceil(h / 32)
    bits = torch.zeros(m, words, dtype=torch.int32, device=dev)
    y = torch.empty(m, h, device=dev)
    K.gemm(m, h, h, K.operand(t, h, K.RC_OP_BN_RELU, (mean, rstd, gam, bet)), K.operand(w2, h), y, h, bias=b2,
           epi=K.RC_EPI_RELU_RES, res=res, ld_res=h, bits_out=bits, ld_bits_out=words)
    assert rel_err(_np(y), _np(y_ref)) < TOL
    got_bits = ((bits.to(torch.int64).unsqueeze(-1) >> torch.arange(32, device=dev)) & 1).reshape(m, -1)[:, :h].bool()
    near_zero = o_ref.abs() < 1e-5
    assert bool(((got_bits == (o_ref > 0)) | near_zero).all())


@pytest.mark.parametrize("m,h", [(20000, 128), (16385, 256)])
def test_tc_node_mlp_backward_chain(dev, m, h):
    """d z = (bitmask(dy) @ W2) masked by BN(t) > 0 with the two BatchNorm column sums; then the BatchNorm-backward
    prologue (d t = c0 dz + c1 (t - mean) + c2) feeding the data-gradient and both weight gradients."""
    from raincast_gnn_b200 import _lib, kernels as K
    g = torch.Generator(device=dev).manual_seed(3 * m + h)
    t = torch.randn(m, h, generator=g, device=dev) * 1.5 + 0.3
    agg = torch.randn(m, h, generator=g, device=dev)
    dy = torch.randn(m, h, generator=g, device=dev)
    w1 = torch.randn(h, h, generator=g, device=dev) / math.sqrt(h)
    w2 = torch.randn(h, h, generator=g, device=dev) / math.sqrt(h)
    gam = torch.rand(h, generator=g, device=dev) + 0.5
    bet = torch.randn(h, generator=g, device=dev) * 0.3
    mask_bits = torch.randint(0, 2, (m, h), generator=g, device=dev).bool()
    words = math.ceil(h / 32)
    packed = (mask_bits.reshape(m, words, 32).to(torch.int64) << torch.arange(32, device=dev)).sum(-1)
    bits = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).to(torch.int32)
    mean = t.mean(0)
    rstd = 1 / torch.sqrt(t.var(0, unbiased=False) + 1e-5)
    # float64 restatement
    t64, mean64, rstd64 = t.double(), mean.double(), rstd.double()
    hat = (t64 - mean64) * rstd64
    z = gam.double() * hat + bet.double()
    do = dy.double() * mask_bits
    dz_ref = (do @ w2.double()) * (z > 0)
    near = z.abs() < 1e-6
    s0_ref, s1_ref = dz_ref.sum(0), (dz_ref * hat).sum(0)
    row_tile = K.gemm_row_tile(m, h, h)
    tiles = math.ceil(m / row_tile)
    stats = torch.empty(tiles, 2, h, device=dev)
    dz = torch.empty(m, h, device=dev)
    do_op = K.operand(dy, h, K.RC_OP_BITMASK, bits=bits, ld_bits=words)
    K.gemm(m, h, h, do_op, K.operand(w2, h), dz, h, b_layout=K.RC_B_RED, epi=K.RC_EPI_BN_RELU_BWD, e_aux=t, ld_e_aux=h,
           e_p=(mean, rstd, gam, bet), stats=stats)
    ok = ~near
    assert rel_err(_np(dz * ok), _np(dz_ref * ok)) < TOL
    if not bool(near.any()):
        assert rel_err(_np(stats[:, 0].double().sum(0)), _np(s0_ref)) < TOL
        assert rel_err(_np(stats[:, 1].double().sum(0)), _np(s1_ref)) < TOL
    c0 = torch.randn(h, generator=g, device=dev)
    c1 = torch.randn(h, generator=g, device=dev) * 0.1
    c2 = torch.randn(h, generator=g, device=dev) * 0.1
    dt_ref = c0.double() * dz.double() + c1.double() * (t64 - mean64) + c2.double()
    dt_op = K.operand(dz, h, K.RC_OP_AFFINE2, (c0, c1, c2, mean), aux=t, ld_aux=h)
    d_agg = torch.empty(m, h, device=dev)
    K.gemm(m, h, h, dt_op, K.operand(w1, h), d_agg, h, b_layout=K.RC_B_RED)
    assert rel_err(_np(d_agg), _np(dt_ref @ w1.double())) < TOL
    # weight gradients: d W1 = d t^T agg (+ bias gradient), d W2 = d o^T u with u = relu(BN(t)) recomputed
    G = {"w": torch.empty(h, h, device=dev), "b": torch.empty(h, device=dev)}
    sink = K.GradSink(dev)
    K.linear_bwd_weight(dt_op, K.operand(agg, h), m, h, h, G["w"], G["b"], sink)
    sink.flush()
    assert rel_err(_np(G["w"]), _np(dt_ref.T @ agg.double())) < TOL
    assert rel_err(_np(G["b"]), _np(dt_ref.sum(0))) < TOL
    u_ref = torch.relu(z)
    sink = K.GradSink(dev)
    K.linear_bwd_weight(do_op, K.operand(t, h, K.RC_OP_BN_RELU, (mean, rstd, gam, bet)), m, h, h, G["w"], G["b"], sink)
    sink.flush()
    assert rel_err(_np(G["w"]), _np(do.T @ u_ref)) < TOL
    assert rel_err(_np(G["b"]), _np(do.sum(0))) < TOL


@pytest.mark.parametrize("m,n,k", [(20000, 128, 35), (16384, 128, 128), (30000, 128, 163), (17001, 512, 512), (16400, 40, 128)])
def test_tc_weight_grad(dev, m, n, k):
    """dw[n, k] = dy[m, n]^T @ x[m, k] over >= 16384 samples, with the bias gradient."""
    from raincast_gnn_b200 import _lib, kernels as K
    g = torch.Generator(device=dev).manual_seed(11 + m + n + k)
    dy = torch.randn(m, n, generator=g, device=dev)
    x = torch.randn(m, k, generator=g, device=dev)
    assert int(_lib.lib().rc_gemm_tc_wgrad_splits(n, k, m)) > 0
    dw, db = torch.empty(n, k, device=dev), torch.empty(n, device=dev)
    sink = K.GradSink(dev)
    K.linear_bwd_weight(K.operand(dy, n), K.operand(x, k), m, n, k, dw, db, sink)
    sink.flush()
    assert rel_err(_np(dw), _np(dy.double().T @ x.double())) < TOL
    assert rel_err(_np(db), _np(dy.double().sum(0))) < TOL
