"""GPU parity of the individual kernels (through the C ABI) against the CPU oracle / plain fp32 torch
restatements of the same op.  Tolerance: max|a-b| / max|b| <= 1e-5 (fp32, BASELINE.json north_star);
integer outputs bit-exact."""
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
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------ graph
@pytest.mark.parametrize("name,batch", [("ref122_d100", 8), ("ref122_d1", 3), ("asym23", 5), ("n7_d300", 1)])
def test_csr_device_bit_exact(dev, golden_graph, name, batch):
    from oracle import graph as og
    from raincast_gnn_b200 import graph as G
    ei, ea = golden_graph[f"{name}.edge_index"], golden_graph[f"{name}.edge_attr"]
    n = int(ei.max()) + 1
    ei_b, ea_b = og.collate_edges(ei, ea, n, batch)
    want = og.csr_layout(ei_b, ea_b, n * batch)
    sg = G.build_station_graph(torch.from_numpy(ei_b).to(dev), torch.from_numpy(ea_b).to(dev), n * batch)
    for k, v in want.items():
        got = _np(getattr(sg, k))
        assert got.dtype == v.dtype, k
        assert np.array_equal(got.view(np.uint32), v.view(np.uint32)), k
    assert torch.equal(sg.edge_index().cpu(), torch.from_numpy(ei_b))          # Appendix B round trip
    host = G.build_station_graph(torch.from_numpy(ei_b), torch.from_numpy(ea_b), n * batch)
    for k in want:
        assert torch.equal(getattr(host, k), getattr(sg, k).cpu()), k


def test_csr_device_empty_and_invalid(dev):
    from raincast_gnn_b200 import _lib, graph as G
    sg = G.build_station_graph(torch.zeros((2, 0), dtype=torch.long, device=dev), torch.zeros((0, 1), device=dev), 4)
    assert sg.rowptr.tolist() == [0, 0, 0, 0, 0] and sg.col.numel() == 0
    bad = torch.tensor([[0, 5], [1, 1]], dtype=torch.long, device=dev)
    with pytest.raises(_lib.RcError):
        G.build_station_graph(bad, torch.ones(2, device=dev), 4)
    with pytest.raises(_lib.RcError):
        G.build_station_graph(bad.cpu(), torch.ones(2), 4)


def test_csr_large_graph_properties(dev):
    """Config-4 sized graph: size-independent properties (round trip, rev o rev = id, degree sums)."""
    from raincast_gnn_b200 import graph as G
    from raincast_gnn_b200.utils import synthetic as syn
    n = 100_000
    ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
    assert ei.shape[1] == 2_978_560
    sg = G.build_station_graph(ei.to(dev), ea.to(dev), n)
    assert torch.equal(sg.edge_index().cpu(), ei)
    assert int(sg.rowptr[-1]) == ei.shape[1] and int(sg.t_rowptr[-1]) == ei.shape[1]
    rev = sg.rev.long()
    assert bool((rev >= 0).all()) and torch.equal(rev[rev], torch.arange(ei.shape[1], device=dev))
    assert torch.equal(sg.attr[rev], sg.attr)                                 # symmetric distances
    host = G.build_station_graph(ei, ea, n)
    for k in ("rowptr", "col", "perm", "t_rowptr", "t_dst", "t_slot", "rev"):
        assert torch.equal(getattr(host, k), getattr(sg, k).cpu()), k


# ------------------------------------------------------------------------------------------------ GEMM family
def _run_gemm(dev, **kw):
    from raincast_gnn_b200 import kernels as K
    return K.gemm(**kw)


@pytest.mark.parametrize("m,n,k", [(976, 128, 128), (30, 32, 7), (257, 5, 128), (1000, 163, 35), (64, 256, 512)])
@pytest.mark.parametrize("rm", [0, 1, 2, 4, 8])
def test_gemm_forward_layout(dev, m, n, k, rm):
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator().manual_seed(m + n + k)
    x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g), torch.randn(n, generator=g)
    want = torch.relu(x.double() @ w.double().T + 2.0 * b.double())
    y = torch.empty(m, n, device=dev)
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)          # keep the device tensors alive: operand() holds raw pointers
    K.gemm(m, n, k, K.operand(xd, k), K.operand(wd, k), y, n, bias=bd, bias_scale=2.0,
           epi=K.RC_EPI_RELU, rows_per_warp=rm)
    assert rel_err(_np(y), want.numpy()) < TOL


@pytest.mark.parametrize("m,n,k", [(976, 128, 128), (30, 7, 32), (257, 128, 5), (100, 35, 163)])
@pytest.mark.parametrize("rm", [0, 1, 8])
def test_gemm_backward_data_layout(dev, m, n, k, rm):
    """dx[m, n] = dy[m, k] @ w[k, n]   (w stored [k][n]: RC_B_RED), masked by aux > 0."""
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator().manual_seed(7 + m + n + k)
    dy, w, aux = torch.randn(m, k, generator=g), torch.randn(k, n, generator=g), torch.randn(m, n, generator=g)
    want = (dy.double() @ w.double()) * (aux > 0)
    out = torch.empty(m, n, device=dev)
    dyd, wd, auxd = dy.to(dev), w.to(dev), aux.to(dev)
    K.gemm(m, n, k, K.operand(dyd, k), K.operand(wd, n), out, n, b_layout=K.RC_B_RED,
           epi=K.RC_EPI_MASK_POS, e_aux=auxd, ld_e_aux=n, rows_per_warp=rm)
    assert rel_err(_np(out), want.numpy()) < TOL


@pytest.mark.parametrize("m,n,k,splits", [(976, 128, 128, 15), (976, 128, 35, 7), (50, 5, 128, 1), (3000, 32, 32, 4), (130, 128, 163, 3)])
@pytest.mark.parametrize("rm", [0, 1, 2, 4, 8])
def test_gemm_weight_grad_layout(dev, m, n, k, splits, rm):
    """dw[n, k] = dy[m, n]^T @ x[m, k] with a split reduction and the bias gradient (column sums of dy)."""
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator().manual_seed(11 + m + n + k)
    dy, x = torch.randn(m, n, generator=g), torch.randn(m, k, generator=g)
    part = torch.empty(splits, n, k, device=dev)
    cs = torch.empty(splits, n, device=dev)
    dyd, xd = dy.to(dev), x.to(dev)
    K.gemm(n, k, m, K.operand(dyd, n), K.operand(xd, k), part, k, a_layout=K.RC_A_RED, b_layout=K.RC_B_RED,
           splits=splits, split_stride=n * k, colsum_a=cs, rows_per_warp=rm)
    assert rel_err(_np(part.double().sum(0)), (dy.double().T @ x.double()).numpy()) < TOL
    assert rel_err(_np(cs.double().sum(0)), dy.double().sum(0).numpy()) < TOL


def test_gemm_two_segments(dev):
    """Linear(cat([x, emb])) as two reduction segments (models/gnn.py:134-135)."""
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator().manual_seed(5)
    m, f, h = 976, 35, 128
    x, emb = torch.randn(m, f, generator=g), torch.randn(m, h, generator=g)
    w, b = torch.randn(h, f + h, generator=g) * 0.1, torch.randn(h, generator=g)
    P = {"dimred_w": w.to(dev), "dimred_b": b.to(dev)}
    y, _ = K.dimred_fwd(P, x.to(dev), emb.to(dev))
    want = torch.cat([x, emb], 1).double() @ w.double().T + b.double()
    assert rel_err(_np(y), want.numpy()) < TOL


@pytest.mark.parametrize("m,h", [(976, 128), (61, 32), (1000, 256)])
def test_gemm_bn_stats_and_prologue(dev, m, h):
    """Linear -> BatchNorm1d(train) -> ReLU -> Linear (+ReLU, residual, bit mask) vs torch modules."""
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator().manual_seed(m + h)
    x = torch.randn(m, h, generator=g) + 3.0                      # non-zero mean: exercises the centred variance
    lin1, bn, lin2 = torch.nn.Linear(h, h), torch.nn.BatchNorm1d(h), torch.nn.Linear(h, h)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
    res = torch.randn(m, h, generator=g)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    bn.train()
    with torch.no_grad():
        t_ref = lin1(x)
        o_ref = lin2(torch.relu(bn(t_ref)))
        y_ref = res + torch.relu(o_ref)
    t = torch.empty(m, h, device=dev)
    row_tile = K.gemm_row_tile(m, h)
    tiles = math.ceil(m / row_tile)
    stats = torch.empty(tiles, 2, h, device=dev)
    xd = x.to(dev)
    w1, b1, w2, b2 = (p.detach().to(dev) for p in (lin1.weight, lin1.bias, lin2.weight, lin2.bias))
    gam, bet = bn.weight.detach().to(dev), bn.bias.detach().to(dev)
    K.gemm(m, h, h, K.operand(xd, h), K.operand(w1, h), t, h, bias=b1, epi=K.RC_EPI_BN_STATS, stats=stats)
    mean, rstd = torch.empty(h, device=dev), torch.empty(h, device=dev)
    rm, rv, nbt = rm0.to(dev), rv0.to(dev), torch.zeros((), dtype=torch.long, device=dev)
    from raincast_gnn_b200 import _lib
    _lib.check(_lib.lib().rc_bn_stats_finalize(stats.data_ptr(), tiles, row_tile, m, h, 1e-5, 0.1, mean.data_ptr(), rstd.data_ptr(),
                                               rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert rel_err(_np(t), t_ref.numpy()) < TOL
    assert rel_err(_np(mean), t_ref.double().mean(0).numpy()) < TOL
    assert rel_err(_np(rstd), (1 / torch.sqrt(t_ref.double().var(0, unbiased=False) + 1e-5)).numpy()) < TOL
    assert rel_err(_np(rm), bn.running_mean.numpy()) < TOL and rel_err(_np(rv), bn.running_var.numpy()) < TOL
    assert int(nbt) == 1
    words = math.ceil(h / 32)
    bits = torch.empty(m, words, dtype=torch.int32, device=dev)
    y = torch.empty(m, h, device=dev)
    resd = res.to(dev)
    K.gemm(m, h, h, K.operand(t, h, K.RC_OP_BN_RELU, (mean, rstd, gam, bet)), K.operand(w2, h), y, h, bias=b2,
           epi=K.RC_EPI_RELU_RES, res=resd, ld_res=h, bits_out=bits, ld_bits_out=words)
    assert rel_err(_np(y), y_ref.numpy()) < TOL
    got_bits = ((bits.cpu().to(torch.int64).unsqueeze(-1) >> torch.arange(32)) & 1).reshape(m, -1)[:, :h].bool()
    want_bits = o_ref > 0
    near_zero = o_ref.abs() < 1e-5                                   # sign may differ only at rounding level
    assert bool(((got_bits == want_bits) | near_zero).all())


# ------------------------------------------------------------------------------------------------ GINE aggregation
def _gine_ref(x, ei, ea, w, b, eps):
    src, dst = ei[0], ei[1]
    msg = torch.relu(x[src] + ea * w[None, :] + b[None, :])
    return torch.zeros_like(x).index_add_(0, dst, msg) + (1 + eps) * x


@pytest.mark.parametrize("name,batch,h", [("ref122_d100", 8, 128), ("asym23", 3, 32), ("ref122_d1", 2, 64), ("n40_d150", 4, 256)])
def test_gine_aggregation_fwd_bwd(dev, golden_graph, name, batch, h):
    from oracle import graph as og
    from raincast_gnn_b200 import _lib, graph as G
    ei, ea = golden_graph[f"{name}.edge_index"], golden_graph[f"{name}.edge_attr"]
    n = int(ei.max()) + 1
    ei_b, ea_b = og.collate_edges(ei, ea, n, batch)
    m = n * batch
    g = torch.Generator().manual_seed(h + m)
    x = torch.randn(m, h, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(h, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(h, generator=g, dtype=torch.float64, requires_grad=True)
    eps = torch.tensor([0.3], dtype=torch.float64, requires_grad=True)
    gout = torch.randn(m, h, generator=g, dtype=torch.float64)
    add = torch.randn(m, h, generator=g, dtype=torch.float64)
    want = _gine_ref(x, torch.from_numpy(ei_b), torch.from_numpy(ea_b).double(), w, b, eps)
    want.backward(gout)
    sg = G.build_station_graph(torch.from_numpy(ei_b), torch.from_numpy(ea_b), m).to(dev)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    xd, wd, bd, ed, gd, ad = (t.detach().float().to(dev) for t in (x, w, b, eps, gout, add))
    hh = torch.empty(m, h, device=dev)
    _lib.check(L.rc_gine_aggr_fwd(xd.data_ptr(), sg.rowptr.data_ptr(), sg.col.data_ptr(), sg.attr.data_ptr(), wd.data_ptr(),
                                  bd.data_ptr(), ed.data_ptr(), hh.data_ptr(), m, h, st))
    assert rel_err(_np(hh), want.detach().numpy()) < TOL
    nb = L.rc_gine_aggr_bwd_nblocks(m, h)
    part = torch.empty(nb, 3, h, device=dev)
    dx = torch.empty(m, h, device=dev)
    _lib.check(L.rc_gine_aggr_bwd(gd.data_ptr(), xd.data_ptr(), sg.t_rowptr.data_ptr(), sg.t_dst.data_ptr(), sg.t_attr.data_ptr(),
                                  wd.data_ptr(), bd.data_ptr(), ed.data_ptr(), ad.data_ptr(), dx.data_ptr(), part.data_ptr(), m, h, st))
    dw, db, de = torch.empty(h, device=dev), torch.empty(h, device=dev), torch.empty(1, device=dev)
    _lib.check(L.rc_gine_aggr_bwd_finalize(part.data_ptr(), nb, h, dw.data_ptr(), db.data_ptr(), de.data_ptr(), st))
    assert rel_err(_np(dx), (x.grad + add).numpy()) < TOL
    assert rel_err(_np(dw), w.grad.numpy()) < TOL
    assert rel_err(_np(db), b.grad.numpy()) < TOL
    assert rel_err(_np(de), eps.grad.numpy()) < TOL


@pytest.mark.parametrize("name,batch,h,hid", [("ref122_d100", 8, 128, 128), ("asym23", 3, 32, 64), ("n40_d150", 4, 256, 128)])
def test_gine_aggregation_as_gemm_prologue_is_bit_identical(dev, golden_graph, name, batch, h, hid):
    """N1: the aggregation fused into the layer's first Linear (RC_OP_GINE_AGGR A-operand prologue + a_out) against the
    stand-alone aggregation kernel followed by the same GEMM: identical bits in the aggregate, the Linear output and the
    BatchNorm tile statistics."""
    from oracle import graph as og
    from raincast_gnn_b200 import graph as G, kernels as K
    ei, ea = golden_graph[f"{name}.edge_index"], golden_graph[f"{name}.edge_attr"]
    n = int(ei.max()) + 1
    ei_b, ea_b = og.collate_edges(ei, ea, n, batch)
    m = n * batch
    sg = G.build_station_graph(torch.from_numpy(ei_b), torch.from_numpy(ea_b), m).to(dev)
    g = torch.Generator().manual_seed(h + m)
    x = torch.randn(m, h, generator=g).to(dev)
    w_e, b_e = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev)
    eps = torch.tensor([0.3], device=dev)
    w0 = (torch.randn(hid, h, generator=g) / h ** 0.5).to(dev)
    b0 = torch.randn(hid, generator=g).to(dev)
    # two kernels
    agg = torch.empty(m, h, device=dev)
    K.gine_aggr_fwd(x, sg, w_e, b_e, eps, agg, tiled=False)
    row_tile = K.gemm_row_tile(m, hid, h)
    tiles = -(-m // row_tile)
    t_two, st_two = torch.empty(m, hid, device=dev), torch.empty(tiles, 2, hid, device=dev)
    K.gemm(m, hid, h, K.operand(agg, h), K.operand(w0, h), t_two, hid, bias=b0, epi=K.RC_EPI_BN_STATS, stats=st_two)
    # one kernel
    agg_one = torch.full((m, h), float("nan"), device=dev)
    t_one, st_one = torch.empty(m, hid, device=dev), torch.empty(tiles, 2, hid, device=dev)
    a_op = K.operand(x, h, K.RC_OP_GINE_AGGR, (w_e, b_e, eps, None), aux=sg.attr, idx0=sg.rowptr, idx1=sg.col)
    K.gemm(m, hid, h, a_op, K.operand(w0, h), t_one, hid, bias=b0, epi=K.RC_EPI_BN_STATS, stats=st_one, a_out=agg_one, ld_a_out=h)
    torch.cuda.synchronize()
    assert torch.equal(agg_one, agg)
    assert torch.equal(t_one, t_two)
    assert torch.equal(st_one, st_two)


# ------------------------------------------------------------------------------------------------ DeepSets
@pytest.mark.parametrize("m,em,f,h", [(976, 11, 35, 128), (30, 4, 7, 32), (100, 51, 35, 128), (64, 10, 35, 256), (17, 3, 5, 32),
                                       (50, 6, 64, 128), (333, 7, 20, 200)])
def test_deepsets_block(dev, m, em, f, h):
    from oracle.model import DeepSetEncoder
    from raincast_gnn_b200 import kernels as K
    torch.manual_seed(m + em)
    enc = DeepSetEncoder(f, h, h).double()
    ens = torch.randn(m, em, f, dtype=torch.float64)
    gout = torch.randn(m, h, dtype=torch.float64)
    emb = enc(ens)
    emb.backward(gout)
    names = {"phi0_w": enc.phi[0].weight, "phi0_b": enc.phi[0].bias, "phi2_w": enc.phi[2].weight, "phi2_b": enc.phi[2].bias,
             "rho0_w": enc.rho[0].weight, "rho0_b": enc.rho[0].bias, "rho2_w": enc.rho[2].weight, "rho2_b": enc.rho[2].bias}
    P = {k: v.detach().float().to(dev).contiguous() for k, v in names.items()}
    got, saved = K.deepsets_fwd(P, ens.float().to(dev))
    assert rel_err(_np(got), emb.detach().numpy()) < TOL
    G = {k: torch.empty_like(v) for k, v in P.items()}
    K.deepsets_bwd(P, saved, gout.float().to(dev), G)
    for k, v in names.items():
        assert rel_err(_np(G[k]), v.grad.numpy()) < TOL, k


# ------------------------------------------------------------------------------------------------ links + CRPS
CRPS_CFG = [("mixed_u", "MixedLoss", "True", 5, 3), ("mixed", "MixedLoss", "False", 4, 2),
            ("mixednormal", "MixedNormalCRPS", "False", 3, 1), ("normal", "NormalCRPS", "False", 2, 0)]


@pytest.mark.parametrize("tag,loss,grad_u,width,kind", CRPS_CFG)
@pytest.mark.parametrize("seed,n", [(11, 257), (12, 64)])
def test_crps_modules_vs_reference_fixtures(dev, golden_crps, tag, loss, grad_u, width, kind, seed, n):
    """PostProcess + loss_fn.crps through the public module API against the reference's own outputs."""
    from oracle.make_golden import crps_case_inputs
    from raincast_gnn_b200.models import MixedLoss, MixedNormalCRPS, NormalCRPS, PostProcess
    raw, y = crps_case_inputs(seed, n, width)
    key = f"{tag}.s{seed}"
    raw_d = raw.to(dev).requires_grad_(True)
    post = PostProcess(loss, grad_u)(raw_d)
    post.retain_grad()
    fn = {"MixedLoss": lambda: MixedLoss(grad_u=(grad_u == "True"), xi=0.5, u=None if grad_u == "True" else 1.71),
          "MixedNormalCRPS": MixedNormalCRPS, "NormalCRPS": NormalCRPS}[loss]()
    val = fn.crps(post, y.to(dev))
    val.backward()
    want = float(golden_crps[f"{key}.loss"])
    assert val.dtype == torch.float64
    assert rel_err(_np(post), golden_crps[f"{key}.post"]) < 1e-6
    assert abs(val.item() - want) <= TOL * abs(want)
    assert torch.isfinite(raw_d.grad).all()
    assert rel_err(_np(raw_d.grad), golden_crps[f"{key}.draw"]) < TOL
    sane = golden_crps[f"{key}.post"][:, 1] > 1e-2
    if width >= 4:
        sane &= golden_crps[f"{key}.post"][:, 3] > 1e-3
    assert rel_err(_np(post.grad)[sane], golden_crps[f"{key}.dpost"][sane]) < 3e-5     # see tests/test_crps_math_host.py
    # raw mode (links fused) must agree with the two-kernel path
    from raincast_gnn_b200 import kernels as K
    loss2, d_raw2, nv = K.crps_fwd_bwd(raw.to(dev), y.to(dev), kind, raw_input=True, u=1.71, xi=0.5, t=5.0)
    assert abs(loss2.item() - val.item()) < 1e-6 * abs(want)
    assert rel_err(_np(d_raw2), _np(raw_d.grad)) < 1e-6
    assert int(nv) == int((~torch.isnan(y)).sum())


def test_crps_large_and_edge_cases(dev):
    """Full-size property checks: all-NaN targets give NaN loss and zero gradient; a big batch agrees with
    the oracle on a random sample of rows (per-row losses are independent)."""
    from oracle import losses as ol
    from raincast_gnn_b200 import kernels as K
    from raincast_gnn_b200.utils import synthetic as syn
    m = 1 << 20
    g = torch.Generator().manual_seed(3)
    raw = torch.randn(m, 5, generator=g)
    y = syn.log_precip_targets(m, seed=9)
    loss, d_raw, nv = K.crps_fwd_bwd(raw.to(dev), y.to(dev), 3, raw_input=True)
    idx = torch.randperm(m, generator=g)[:4096]
    r = raw[idx].clone().requires_grad_(True)
    per = ol.mixed_loss_crps(ol.postprocess(r, "MixedLoss", "True"), y[idx], grad_u=True, xi=0.5, reduce=False)
    per.sum().backward()
    ok = ~torch.isnan(y[idx])
    want = r.grad[ok] / int(nv)
    assert rel_err(_np(d_raw[idx.to(dev)])[ok.numpy()], want.numpy()) < TOL
    full = ol.mixed_loss_crps(ol.postprocess(raw, "MixedLoss", "True"), y, grad_u=True, xi=0.5)
    assert abs(loss.item() - full.item()) < TOL * abs(full.item())
    ynan = torch.full((100,), float("nan"))
    loss, d, nv = K.crps_fwd_bwd(raw[:100].to(dev), ynan.to(dev), 3, raw_input=True)
    assert math.isnan(loss.item()) and int(nv) == 0 and float(d.abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ AdamW
def test_adamw_kernel_matches_torch(dev):
    from raincast_gnn_b200 import _lib
    torch.manual_seed(0)
    n = 209_929
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-4)
    p = p0.to(dev)
    m1, v, step = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros((), dtype=torch.long, device=dev)
    L = _lib.lib()
    for it in range(5):
        gr = torch.randn(n) * (0.1 + it)
        ref.grad = gr.clone()
        opt.step()
        grd = gr.to(dev)
        _lib.check(L.rc_adamw_step(p.data_ptr(), grd.data_ptr(), m1.data_ptr(), v.data_ptr(), step.data_ptr(), n, 1e-4, 0.9, 0.999,
                                   1e-8, 0.01, 1.0, torch.cuda.current_stream().cuda_stream))
    assert int(step) == 5
    assert rel_err(_np(p), ref.detach().numpy()) < 1e-6
    # the update itself (~5e-4 on parameters of size ~1): limited by fp32 rounding of p, not by the kernel
    assert rel_err(_np(p - p0.to(dev)), (ref.detach() - p0).numpy()) < 2e-3


@pytest.mark.parametrize("h", [128, 256])
def test_gine_aggregation_large_graph_path(dev, h):
    """The contiguous-range kernels used from 16 384 rows up, on a 20k-node radius graph.  54 M ReLU units: inputs
    are dyadic rationals so that every pre-activation x_j + a*w + b is exact in fp32 and float64 alike and no unit
    sits within rounding of its threshold (the comparison then tests the kernel, not the conditioning)."""
    from raincast_gnn_b200 import _lib, graph as G
    from raincast_gnn_b200.utils import synthetic as syn
    m = 20_000
    coords = syn.station_coords(m, 450.0, seed=2)
    ei, _ = G.radius_graph_from_coords(coords, syn.scaled_graph_radius(m, 450.0, 20.0))
    g = torch.Generator().manual_seed(h)

    def dyadic(*shape, scale=8.0):
        return torch.round(torch.randn(*shape, generator=g, dtype=torch.float64) * scale) / scale
    ea = (torch.randint(1, 64, (ei.shape[1], 1), generator=g).double() / 4.0)
    x, w, b = dyadic(m, h).requires_grad_(True), dyadic(h).requires_grad_(True), dyadic(h).requires_grad_(True)
    eps = torch.tensor([-0.25], dtype=torch.float64, requires_grad=True)
    gout = torch.randn(m, h, generator=g, dtype=torch.float64)
    want = _gine_ref(x, ei, ea, w, b, eps)
    want.backward(gout)
    sg = G.build_station_graph(ei, ea.float(), m).to(dev)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    xd, wd, bd, ed, gd = (t.detach().float().to(dev) for t in (x, w, b, eps, gout))
    hh = torch.empty(m, h, device=dev)
    _lib.check(L.rc_gine_aggr_fwd(xd.data_ptr(), sg.rowptr.data_ptr(), sg.col.data_ptr(), sg.attr.data_ptr(), wd.data_ptr(),
                                  bd.data_ptr(), ed.data_ptr(), hh.data_ptr(), m, h, st))
    assert rel_err(_np(hh), want.detach().numpy()) < TOL
    nb = L.rc_gine_aggr_bwd_nblocks(m, h)
    part = torch.empty(nb, 3, h, device=dev)
    dx = torch.empty(m, h, device=dev)
    _lib.check(L.rc_gine_aggr_bwd(gd.data_ptr(), xd.data_ptr(), sg.t_rowptr.data_ptr(), sg.t_dst.data_ptr(), sg.t_attr.data_ptr(),
                                  wd.data_ptr(), bd.data_ptr(), ed.data_ptr(), None, dx.data_ptr(), part.data_ptr(), m, h, st))
    dw, db, de = torch.empty(h, device=dev), torch.empty(h, device=dev), torch.empty(1, device=dev)
    _lib.check(L.rc_gine_aggr_bwd_finalize(part.data_ptr(), nb, h, dw.data_ptr(), db.data_ptr(), de.data_ptr(), st))
    assert rel_err(_np(dx), x.grad.numpy()) < TOL
    assert rel_err(_np(dw), w.grad.numpy()) < TOL and rel_err(_np(db), b.grad.numpy()) < TOL
    assert rel_err(_np(de), eps.grad.numpy()) < TOL


@pytest.mark.parametrize("m,em,f,h", [(2000, 51, 35, 128), (6500, 11, 35, 128), (1500, 51, 35, 256), (3000, 30, 20, 128)])
def test_deepsets_tensor_core_path_fp32(dev, m, em, f, h):
    """>= 65 536 member rows: rc_deepsets_pool_fwd runs on tcgen05 with the 3xTF32 split; still held to 1e-5."""
    from raincast_gnn_b200 import _lib
    g = torch.Generator().manual_seed(m + em)
    ens = torch.randn(m, em, f, generator=g)
    w1 = (torch.rand(h, f, generator=g) * 2 - 1) / f ** 0.5
    b1 = torch.randn(h, generator=g) * 0.1
    want = torch.relu(ens.double() @ w1.double().T + b1.double()).sum(1)
    ed, wd, bd = ens.to(dev), w1.to(dev), b1.to(dev)
    out = torch.empty(m, h, device=dev)
    _lib.check(_lib.lib().rc_deepsets_pool_fwd(ed.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), m, em, f, h,
                                               torch.cuda.current_stream().cuda_stream))
    assert rel_err(_np(out), want.numpy()) < TOL


@pytest.mark.parametrize("m,em,f,h,shift", [(1301, 51, 35, 128, 1), (6001, 11, 35, 128, 3), (2201, 30, 33, 128, 2)])
def test_deepsets_tensor_core_unaligned_ensemble(dev, m, em, f, h, shift):
    """The member rows travel as 16-byte cp.async chunks of the aligned span around each tile: an ensemble that starts
    1-3 floats past a 16-byte boundary (and ends as far before one) must give the same pooled sums."""
    from raincast_gnn_b200 import _lib
    g = torch.Generator().manual_seed(m + em)
    ens = torch.randn(m, em, f, generator=g)
    w1 = (torch.rand(h, f, generator=g) * 2 - 1) / f ** 0.5
    b1 = torch.randn(h, generator=g) * 0.1
    want = torch.relu(ens.double() @ w1.double().T + b1.double()).sum(1)
    buf = torch.full((m * em * f + shift,), float("nan"), device=dev)       # NaN before the first element: must not be read in
    buf[shift:] = ens.reshape(-1).to(dev)
    ed = buf[shift:]
    assert ed.data_ptr() % 16 == 4 * shift
    wd, bd = w1.to(dev), b1.to(dev)
    out = torch.empty(m, h, device=dev)
    _lib.check(_lib.lib().rc_deepsets_pool_fwd(ed.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), m, em, f, h,
                                               torch.cuda.current_stream().cuda_stream))
    assert rel_err(_np(out), want.numpy()) < TOL


@pytest.mark.parametrize("m,em,f,h", [(2000, 51, 35, 512), (977, 11, 35, 128), (64, 10, 35, 512), (300, 128, 64, 128)])
def test_deepsets_tensor_core_path_bf16(dev, m, em, f, h):
    """BASELINE.json config 5: bf16 operands, fp32 accumulate / pool / output; 1e-2 against the float64 oracle."""
    from raincast_gnn_b200 import _lib
    g = torch.Generator().manual_seed(m + em + h)
    ens = torch.randn(m, em, f, generator=g)
    w1 = (torch.rand(h, f, generator=g) * 2 - 1) / f ** 0.5
    b1 = torch.randn(h, generator=g) * 0.1
    want = torch.relu(ens.double() @ w1.double().T + b1.double()).sum(1)
    ed, wd, bd = ens.to(dev), w1.to(dev), b1.to(dev)
    out = torch.empty(m, h, device=dev)
    _lib.check(_lib.lib().rc_deepsets_pool_fwd_bf16(ed.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), m, em, f, h,
                                                    torch.cuda.current_stream().cuda_stream))
    assert rel_err(_np(out), want.numpy()) < 1e-2
    # and exactly what bf16 operands with fp32 accumulation should give (operands rounded, products exact)
    ref = torch.relu(ens.bfloat16().double() @ w1.bfloat16().double().T + b1.double()).sum(1)
    assert rel_err(_np(out), ref.numpy()) < 1e-5


# ------------------------------------------------------------------------------------------------ station tiles
def _tiled_vs_untiled(dev, sg, m, h, seed, want=None):
    """Runs both aggregation paths on the same inputs; they agree to rounding (the tiled kernels sum class by class and
    add the edge bias once per row, the parameter gradients group their partial sums differently)."""
    from raincast_gnn_b200 import kernels as K
    g = torch.Generator().manual_seed(seed)
    x, gout, add = (torch.randn(m, h, generator=g).to(dev) for _ in range(3))
    w, b = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev)
    eps = torch.tensor([0.2], device=dev)
    L = _lib_mod().lib()
    outs = {}
    for tiled in (False, True):
        hh, dx = torch.empty(m, h, device=dev), torch.empty(m, h, device=dev)
        K.gine_aggr_fwd(x, sg, w, b, eps, hh, tiled=tiled)
        part, nb = K.gine_aggr_bwd(gout, x, sg, w, b, eps, add, dx, tiled=tiled)
        dw, db, de = torch.empty(h, device=dev), torch.empty(h, device=dev), torch.empty(1, device=dev)
        _lib_mod().check(L.rc_gine_aggr_bwd_finalize(part.data_ptr(), nb, h, dw.data_ptr(), db.data_ptr(), de.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream))
        outs[tiled] = (hh, dx, dw, db, de)
    torch.cuda.synchronize()
    assert rel_err(_np(outs[True][0]), _np(outs[False][0])) < 2e-6, "tiled forward differs from the warp-per-row kernel"
    # the two kernels round x_j + a w + b differently (x_j + (a w + b) > 0 vs x_j + a w > -b): a unit within one
    # rounding of its ReLU threshold - a few in 1e8 evaluations - can land on either side and moves one element of dx
    # by g; everything else agrees to rounding
    a, b = _np(outs[True][1]), _np(outs[False][1])
    off = np.abs(a - b) > 2e-6 * np.abs(b).max()
    assert off.sum() <= max(2, 1e-5 * off.size), "tiled backward dx differs from the warp-per-row kernel"
    # each such unit also moves d w_edge by g * a and d b_edge by g
    flip = float(off.sum()) * float(gout.abs().max()) * max(1.0, float(sg.attr.abs().max()))
    for i in (2, 3):
        assert rel_err(_np(outs[True][i]), _np(outs[False][i])) < TOL + flip / float(outs[False][i].abs().max())
    assert rel_err(_np(outs[True][4]), _np(outs[False][4])) < TOL
    return outs[True]


def _lib_mod():
    from raincast_gnn_b200 import _lib
    return _lib


@pytest.mark.parametrize("h", [128, 256, 512])
def test_gine_tiled_radius_graph(dev, h):
    """Station tiles on a 20k-node radius graph (mean degree 20): equal to the untiled kernels to rounding, and within
    1e-5 of the float64 restatement of PyG's GINEConv on dyadic inputs (no unit near its ReLU threshold)."""
    from raincast_gnn_b200 import graph as G, kernels as K
    from raincast_gnn_b200.utils import synthetic as syn
    m = 20_000
    coords = syn.station_coords(m, 450.0, seed=2)
    ei, ea = G.radius_graph_from_coords(coords, syn.scaled_graph_radius(m, 450.0, 20.0))
    sg = G.build_station_graph(ei, ea, m).to(dev)
    tiles = sg.tiles(h)
    assert tiles is not None and tiles[0].n_tiles > 50 and tiles[0].max_staged <= G.tile_limits(h)[0]
    _tiled_vs_untiled(dev, sg, m, h, seed=h)
    # oracle comparison on dyadic data
    g = torch.Generator().manual_seed(h + 1)

    def dyadic(*shape, scale=8.0):
        return torch.round(torch.randn(*shape, generator=g, dtype=torch.float64) * scale) / scale
    ea2 = (torch.randint(1, 64, (ei.shape[1], 1), generator=g).double() / 4.0)
    sg2 = G.build_station_graph(ei, ea2.float(), m).to(dev)
    x, w, b = dyadic(m, h).requires_grad_(True), dyadic(h).requires_grad_(True), dyadic(h).requires_grad_(True)
    eps = torch.tensor([-0.25], dtype=torch.float64, requires_grad=True)
    gout = torch.randn(m, h, generator=g, dtype=torch.float64)
    want = _gine_ref(x, ei, ea2, w, b, eps)
    want.backward(gout)
    xd, wd, bd, ed, gd = (t.detach().float().to(dev) for t in (x, w, b, eps, gout))
    hh, dx = torch.empty(m, h, device=dev), torch.empty(m, h, device=dev)
    K.gine_aggr_fwd(xd, sg2, wd, bd, ed, hh, tiled=True)
    part, nb = K.gine_aggr_bwd(gd, xd, sg2, wd, bd, ed, None, dx, tiled=True)
    dw, db, de = torch.empty(h, device=dev), torch.empty(h, device=dev), torch.empty(1, device=dev)
    _lib_mod().check(_lib_mod().lib().rc_gine_aggr_bwd_finalize(part.data_ptr(), nb, h, dw.data_ptr(), db.data_ptr(), de.data_ptr(),
                                                               torch.cuda.current_stream().cuda_stream))
    assert rel_err(_np(hh), want.detach().numpy()) < TOL
    assert rel_err(_np(dx), x.grad.numpy()) < TOL
    assert rel_err(_np(dw), w.grad.numpy()) < TOL and rel_err(_np(db), b.grad.numpy()) < TOL
    assert rel_err(_np(de), eps.grad.numpy()) < TOL


def test_gine_nan_feature_row(dev):
    """A NaN node feature.  torch.relu(NaN) is NaN, so PyG hands it to every neighbour's sum; both aggregation paths here
    compute the message ReLU with an instruction that drops NaN (fmaxf(NaN, 0) = 0 in the warp-per-row kernels,
    fma.rn.sat in the station-tile kernels): the NaN stays in the row's own self term (1 + eps) x_i and its neighbours
    get the sum of their other messages.  A documented deviation: node features are standardised inputs /
    model activations - only the targets y carry NaN, and the CRPS kernel masks those like the reference."""
    from raincast_gnn_b200 import graph as G, kernels as K
    from raincast_gnn_b200.utils import synthetic as syn
    m, h, bad = 20_000, 128, 4321
    coords = syn.station_coords(m, 450.0, seed=2)
    ei, ea = G.radius_graph_from_coords(coords, syn.scaled_graph_radius(m, 450.0, 20.0))
    sg = G.build_station_graph(ei, ea, m).to(dev)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(m, h, generator=g)
    w, b, eps = torch.randn(h, generator=g), torch.randn(h, generator=g), torch.tensor([0.1])
    wd, bd, ed = (t.to(dev) for t in (w, b, eps))
    neighbours = sorted(set(ei[1][ei[0] == bad].tolist()) - {bad})
    assert len(neighbours) > 5
    # what the neighbours must get: the aggregation over the graph without the edges that leave `bad`
    keep = ei[0] != bad
    sg_cut = G.build_station_graph(ei[:, keep], ea[keep], m).to(dev)
    want = torch.empty(m, h, device=dev)
    K.gine_aggr_fwd(x.to(dev), sg_cut, wd, bd, ed, want, tiled=False)
    x[bad] = float("nan")
    for tiled in (False, True):
        out = torch.empty(m, h, device=dev)
        K.gine_aggr_fwd(x.to(dev), sg, wd, bd, ed, out, tiled=tiled)
        assert torch.isnan(out).any(dim=1).nonzero().flatten().tolist() == [bad], f"tiled={tiled}"
        assert rel_err(_np(out[neighbours]), _np(want[neighbours])) < TOL


def test_gine_tiled_batched_reference_graphs(dev, golden_graph):
    """160 reference graphs in one batch (19 520 stations): one tile per graph, no halo, bitwise equal results."""
    from oracle import graph as og
    from raincast_gnn_b200 import graph as G
    ei, ea = golden_graph["ref122_d100.edge_index"], golden_graph["ref122_d100.edge_attr"]
    batch = 160
    ei_b, ea_b = og.collate_edges(ei, ea, 122, batch)
    m = 122 * batch
    sg = G.build_station_graph(torch.from_numpy(ei_b), torch.from_numpy(ea_b), m).to(dev)
    tiles = sg.tiles(128)
    assert tiles is not None and tiles[0].n_tiles == batch and tiles[0].n_halo == 0 and tiles[1].n_halo == 0
    _tiled_vs_untiled(dev, sg, m, 128, seed=7)


@pytest.mark.parametrize("h,max_src,max_block", [(128, 48, 4096), (256, 64, 1 << 16), (384, 100, 2048)])
def test_gine_tiled_irregular_multigraph(dev, h, max_src, max_block):
    """Tiles forced onto a small asymmetric multigraph (repeated edges, random self loops, rows without edges, small
    staging / block limits so that tiles split and give rows back): same results as the warp-per-row kernels."""
    from raincast_gnn_b200 import graph as G
    g = torch.Generator().manual_seed(h)
    m, e = 900, 9000
    src = torch.randint(0, m - 50, (e,), generator=g)                   # the last 50 nodes gather and feed nothing
    dst = (src + torch.randint(-12, 13, (e,), generator=g)).clamp_(0, m - 51)
    ei = torch.stack([src, dst])
    ei = torch.cat([ei, ei[:, :300]], dim=1)                            # 300 repeated (src, dst) pairs, other attributes
    ea = torch.rand(ei.shape[1], 1, generator=g) * 4.0 - 1.0
    sg = G.build_station_graph(ei, ea, m)
    fwd = G.build_tiles_host(sg.rowptr, sg.col, sg.attr, max_src, max_block, G.TILE_ROW_BYTES)
    bwd = G.build_tiles_host(sg.t_rowptr, sg.t_dst, sg.t_attr, max_src, max_block, G.TILE_ROW_BYTES)
    fwd.verify(sg.rowptr, sg.col, sg.attr)
    bwd.verify(sg.t_rowptr, sg.t_dst, sg.t_attr)
    assert fwd.n_tiles > 8 and fwd.n_entries < ei.shape[1]
    sg = sg.to(dev)
    sg.__dict__["_tiles"] = {"pair": (fwd.to(dev), bwd.to(dev))}
    _tiled_vs_untiled(dev, sg, m, h, seed=3 * h)


# ------------------------------------------------------------------------------------------------ fused head + CRPS
@pytest.mark.parametrize("kind,c", [(3, 5), (2, 4), (1, 3), (0, 2)])
@pytest.mark.parametrize("m,h", [(976, 128), (61, 128), (1003, 256)])
def test_head_crps_fused_matches_the_three_kernel_path(dev, kind, c, m, h):
    """rc_head_crps_fwd_bwd (head Linear + links + CRPS + backward in one launch) against rc_gemm_run + rc_crps_fwd_bwd +
    the head's backward GEMMs, and against float64 for the loss: same valid count, loss, d h, d W, d b."""
    from raincast_gnn_b200 import kernels as K
    from raincast_gnn_b200.utils import synthetic as syn
    g = torch.Generator().manual_seed(100 * kind + m)
    x = torch.randn(m, h, generator=g).to(dev)
    w = (torch.randn(c, h, generator=g) / h ** 0.5).to(dev)
    b = (torch.randn(c, generator=g) * 0.1).to(dev)
    y = syn.log_precip_targets(m, seed=kind + 1).to(dev)
    y[::17] = float("nan")
    P = {"aggr_w": w, "aggr_b": b}
    assert K.head_crps_blocks(m, h) == -(-m // 8)
    assert K.head_crps_blocks(20000, h) == 0 and K.head_crps_blocks(m, 512) == 0
    # three-kernel path
    G1 = {"aggr_w": torch.empty_like(w), "aggr_b": torch.empty_like(b)}
    raw, s_h = K.head_fwd(P, x)
    loss1, d_raw, nv1 = K.crps_fwd_bwd(raw, y, kind, raw_input=True, u=1.71, xi=0.5, t=5.0)
    d1 = K.head_bwd(P, s_h, d_raw, G1)
    # one kernel
    G2 = {"aggr_w": torch.full_like(w, float("nan")), "aggr_b": torch.full_like(b, float("nan"))}
    loss2, d2, nv2 = K.head_crps_fwd_bwd(P, x, y, kind, G2, u=1.71, xi=0.5, t=5.0)
    torch.cuda.synchronize()
    assert int(nv2) == int(nv1) == int((~torch.isnan(y)).sum())
    assert abs(loss2.item() - loss1.item()) < 1e-6 * abs(loss1.item())
    assert rel_err(_np(d2), _np(d1)) < TOL
    assert rel_err(_np(G2["aggr_w"]), _np(G1["aggr_w"])) < TOL
    assert rel_err(_np(G2["aggr_b"]), _np(G1["aggr_b"])) < TOL
    # the loss against float64 links + CRPS of the float64 head output
    from oracle import losses
    raw64 = x.double().cpu() @ w.double().cpu().T + b.double().cpu()
    name, gu = {3: ("MixedLoss", "True"), 2: ("MixedLoss", "False"), 1: ("MixedNormalCRPS", "False"), 0: ("NormalCRPS", "False")}[kind]
    pred = losses.postprocess(raw64, name, gu)
    y64 = y.double().cpu()
    if kind >= 2:
        want = losses.mixed_loss_crps(pred, y64, grad_u=(kind == 3), xi=0.5, u=None if kind == 3 else 1.71)
    elif kind == 1:
        want = losses.mixed_normal_crps(pred, y64)
    else:
        want = losses.normal_crps(pred, y64)
    assert abs(loss2.item() - want.item()) < TOL * abs(want.item())
