"""GPU parity of the tensor-core DeepSets pool BACKWARD (rc_deepsets_pool_bwd on tcgen05: mask recomputation, dh in TMEM,
d W1 = dh^T E as a second contraction) against float64 - with the ReLU mask the kernel used dumped and forced into the
float64 restatement, so that every element is held to 1e-5 (a unit within fp32 rounding of its threshold may
legitimately sit on either side; the dump removes that freedom from the comparison instead of allowing for it)."""
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


def _unpack(bits, h):
    words = bits.to(torch.int64) & 0xFFFFFFFF
    return ((words.unsqueeze(-1) >> torch.arange(32, device=bits.device)) & 1).reshape(bits.shape[0], -1)[:, :h].bool()


@pytest.mark.parametrize("m,em,f,h,bf16", [(7001, 11, 35, 128, 0), (1400, 51, 35, 128, 0), (6500, 11, 35, 256, 0), (1300, 51, 33, 128, 0),
                                           (1400, 51, 35, 512, 1), (6100, 11, 40, 128, 0),
                                           (6500, 11, 35, 64, 0), (6500, 11, 35, 96, 0), (1300, 51, 35, 192, 0)])   # widths that are not multiples of 128
def test_pool_bwd_tensor_cores_mask_matched(dev, m, em, f, h, bf16):
    from raincast_gnn_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(m + em + h)
    ens = torch.randn(m, em, f, generator=g, device=dev)
    w1 = (torch.rand(h, f, generator=g, device=dev) * 2 - 1) / f ** 0.5
    b1 = torch.randn(h, generator=g, device=dev) * 0.1
    dp = torch.randn(m, h, generator=g, device=dev)
    assert m * em >= 65536
    nb = int(L.rc_deepsets_pool_bwd_nblocks(m, em, f, h))
    assert nb == min(148, -(-m // (64 // em))), "the tensor-core path was not selected"
    part = torch.full((nb, h * f + h), float("nan"), device=dev)
    words = (h + 31) // 32
    bits = torch.zeros(m * em, words, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.rc_deepsets_pool_bwd(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), dp.data_ptr(), part.data_ptr(), m, em, f, h, bf16,
                                      bits.data_ptr(), st), "rc_deepsets_pool_bwd")
    torch.cuda.synchronize()
    dw = part[:, :h * f].double().sum(0).reshape(h, f)
    db = part[:, h * f:].double().sum(0)
    e2 = ens.reshape(-1, f)
    if bf16:
        e2, w1r = e2.bfloat16().float(), w1.bfloat16().float()
    else:
        w1r = w1
    pre = e2.double() @ w1r.double().T + b1.double()
    mask = _unpack(bits, h)
    # the dumped mask is the float64 one except where the pre-activation is within rounding of zero
    scale = pre.abs().max().item()
    disagree = mask != (pre > 0)
    assert float(pre[disagree].abs().max() if bool(disagree.any()) else 0.0) < 1e-5 * scale
    assert int(disagree.sum()) < 1e-5 * mask.numel() + 10
    dh = dp.double().repeat_interleave(em, 0) * mask
    assert rel_err(_np(dw), _np(dh.T @ e2.double())) < (1e-2 if bf16 else TOL)
    assert rel_err(_np(db), _np(dh.sum(0))) < (1e-2 if bf16 else TOL)
    # without the dump the result is bit-identical (the dump is read-only instrumentation)
    part2 = torch.empty_like(part)
    _lib.check(L.rc_deepsets_pool_bwd(ens.data_ptr(), w1.data_ptr(), b1.data_ptr(), dp.data_ptr(), part2.data_ptr(), m, em, f, h, bf16,
                                      None, st), "rc_deepsets_pool_bwd")
    torch.cuda.synchronize()
    assert torch.equal(part, part2)
