"""Identical-input parity of every elementwise kernel of the executor against torch float64 ("each
conv/BN kernel's fprop/dgrad/wgrad vs the oracle on identical inputs"): BatchNorm training / eval
forward, BatchNorm backward (with the 1-bit ReLU mask, the shortcut gradient and exact zeros), max-pool forward / backward
with ties, the head combine (bilinear ConvTranspose2d + crop + add) forward / backward, and the residual-epilogue GEMM
the `fast` backward uses.  Reference arithmetic: torchvision resnet.py:143-163,197-204 and
/root/reference/tinyfaces/models/model.py:104-126 restated with torch CPU float64 ops.  Tolerance 1e-5 (max-norm,
relative to the tensor's max), written next to each assert.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _unpack(mask, n):
    """1 bit / element (element i -> bit i % 32 of word i / 32) -> bool [n]."""
    w = mask.cpu().numpy().view(np.uint32)
    bits = ((w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).reshape(-1)
    return torch.from_numpy(bits[:n].copy())


def _bn_inputs(M, C, seed, with_res):
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(M, C, generator=g) * (0.5 + torch.rand(C, generator=g) * 2) + torch.randn(C, generator=g)
    gamma = 0.5 + torch.rand(C, generator=g)
    beta = 0.3 * torch.randn(C, generator=g)
    gamma[1] = 0.0; beta[1] = 0.0                  # a channel whose output is EXACTLY 0 everywhere: ReLU derivative at 0 is 0
    gamma[2] = -0.7                                # negative scale
    res = torch.randn(M, C, generator=g) if with_res else None
    if res is not None:
        res[:, 1] = 0.0
    return y, gamma, beta, res


@pytest.mark.parametrize("M,C,with_res", [(4800, 256, False), (38400 + 37, 64, False), (5001, 1024, True), (777, 512, True),
                                         (19, 128, False)])
def test_bn_train_forward_and_backward_vs_torch_fp64(M, C, with_res):
    from tinyfaces_b200 import ops
    y, gamma, beta, res = _bn_inputs(M, C, M + C, with_res)
    rm, rv = torch.randn(C) * 0.1, torch.rand(C) + 0.5
    d = "cuda:0"
    rm_d, rv_d = rm.to(d), rv.to(d)
    out, mask, mean, rstd = ops.bn_train_fwd(y.to(d), gamma.to(d), beta.to(d), rm_d, rv_d, None if res is None else res.to(d),
                                             relu=True, eps=1e-5, momentum=0.1)
    # ---- forward reference (float64)
    y64 = y.double().requires_grad_(True)
    res64 = None if res is None else res.double().requires_grad_(True)
    mu, var = y64.mean(0), y64.var(0, unbiased=False)
    pre = (y64 - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    if res64 is not None:
        pre = pre + res64
    ref = torch.relu(pre)
    assert _rel(out, ref) < TOL                                                            # 1e-5
    assert _rel(mean, mu) < TOL and _rel(rstd, 1 / torch.sqrt(var + 1e-5)) < TOL           # 1e-5
    assert _rel(rm_d, 0.9 * rm.double() + 0.1 * mu) < TOL                                  # running stats, momentum 0.1
    assert _rel(rv_d, 0.9 * rv.double() + 0.1 * var * M / (M - 1)) < TOL                   # unbiased variance
    bits = _unpack(mask, M * C).reshape(M, C)
    assert torch.equal(bits, out.cpu() > 0)                                                # mask == [out > 0] exactly
    assert not bits[:, 1].any()                                                            # exact zeros -> derivative 0
    # ---- backward on IDENTICAL inputs: the reference gates with the same mask bits (a sign flip of a ~1e-7 pre-activation
    #      between fp32 and fp64 is not a kernel error; out itself is checked above)
    g = torch.Generator().manual_seed(7)
    dout = torch.randn(M, C, generator=g)
    (pre * bits.double() * dout.double()).sum().backward()
    dy, dgamma, dbeta, gout = ops.bn_bwd(dout.to(d), mask, y.to(d), mean, rstd, gamma.to(d), want_g=True)
    gamma64 = gamma.double()
    xhat = (y.double() - mu.detach()) / torch.sqrt(var.detach() + 1e-5)
    gm = dout.double() * bits.double()
    assert _rel(dy, y64.grad) < TOL                                                        # 1e-5
    assert _rel(dgamma, (gm * xhat).sum(0)) < TOL and _rel(dbeta, gm.sum(0)) < TOL         # 1e-5
    assert _rel(gout, gm) == 0.0                                                           # the shortcut's gradient: exact
    if res64 is not None:
        assert _rel(gout, res64.grad) == 0.0
    del gamma64


def test_bn_eval_forward_vs_torch_fp64():
    from tinyfaces_b200 import ops
    M, C = 3001, 256
    y, gamma, beta, res = _bn_inputs(M, C, 5, True)
    rm, rv = torch.randn(C), torch.rand(C) + 0.2
    d = "cuda:0"
    out = ops.bn_eval_fwd(y.to(d), gamma.to(d), beta.to(d), rm.to(d), rv.to(d), res.to(d), relu=True)
    ref = torch.relu((y.double() - rm.double()) / torch.sqrt(rv.double() + 1e-5) * gamma.double() + beta.double() + res.double())
    assert _rel(out, ref) < TOL                                                            # 1e-5
    out2 = ops.bn_eval_fwd(y.to(d), gamma.to(d), beta.to(d), rm.to(d), rv.to(d), None, relu=False, round_tf32=True)
    ref2 = (y.double() - rm.double()) / torch.sqrt(rv.double() + 1e-5) * gamma.double() + beta.double()
    assert _rel(out2, ref2) < 6e-4                                                         # TF32 rounding: 2^-11
    assert torch.equal(out2.view(torch.int32) & 0x1FFF, torch.zeros_like(out2, dtype=torch.int32))     # 13 low bits clear


@pytest.mark.parametrize("B,H,W,C", [(2, 37, 53, 64), (1, 250, 250, 64), (3, 8, 9, 128), (1, 1, 1, 64)])
def test_maxpool_forward_backward_with_ties(B, H, W, C):
    from tinyfaces_b200 import ops
    g = torch.Generator().manual_seed(H * W)
    x = torch.round(torch.randn(B, H, W, C, generator=g) * 2) / 2          # multiples of 0.5: many exact ties per window
    d = "cuda:0"
    out, am = ops.maxpool_fwd(x.to(d))
    xn = x.permute(0, 3, 1, 2).double().contiguous().requires_grad_(True)
    ref = F.max_pool2d(xn, 3, 2, 1)
    assert torch.equal(out.permute(0, 3, 1, 2).double().cpu(), ref.detach())               # exact
    dout = torch.randn(out.shape, generator=g)
    ref.backward(dout.permute(0, 3, 1, 2).double())
    dx = ops.maxpool_bwd(am, dout.to(d), H, W)
    assert _rel(dx.permute(0, 3, 1, 2), xn.grad) < 1e-6                                    # same arg-max on ties; fp32 sums of <= 4 terms


@pytest.mark.parametrize("B,H3,W3,eval_crop", [(2, 63, 63, False), (1, 120, 160, False), (1, 13, 17, True), (8, 12, 16, False)])
def test_head_upsample_crop_add_forward_backward(B, H3, W3, eval_crop):
    """model.py:104-126 on identical inputs; the padded channels (125..127) hold garbage on purpose."""
    from oracle import model_oracle
    from tinyfaces_b200 import ops
    Cn, Cp = 125, 128
    H4, W4 = (H3 - 1) // 2 + 1, (W3 - 1) // 2 + 1
    g = torch.Generator().manual_seed(B * H3 + W3)
    s3 = torch.randn(B, H3, W3, Cp, generator=g)
    s4 = torch.randn(B, H4, W4, Cp, generator=g)
    up_w = model_oracle.bilinear_upsample_weight(Cn)
    d = "cuda:0"
    out = ops.head_upsample_add_fwd(s3.to(d), s4.to(d), up_w.to(d), Cn)
    a3 = s3[..., :Cn].permute(0, 3, 1, 2).double().contiguous().requires_grad_(True)
    a4 = s4[..., :Cn].permute(0, 3, 1, 2).double().contiguous().requires_grad_(True)
    score4 = F.conv_transpose2d(a4, up_w.double(), stride=2, padding=1)
    if eval_crop:                                                                          # model.py:110-121
        cv, cu = score4.size(2) - H3, score4.size(3) - W3
        cv = -score4.size(2) if cv == 0 else cv
        cu = -score4.size(3) if cu == 0 else cu
        score4 = score4[:, :, 0:-cv, 0:-cu]
    else:                                                                                  # model.py:122-124
        score4 = score4[:, :, 0:H3, 0:W3]
    ref = a3 + score4
    assert _rel(out, ref) < TOL                                                            # 1e-5
    dout = torch.randn(out.shape, generator=g)
    ref.backward(dout.double())
    ds3, ds4 = ops.head_upsample_add_bwd(dout.to(d), up_w.to(d), H4, W4, Cp)
    assert _rel(ds3[..., :Cn].permute(0, 3, 1, 2), a3.grad) == 0.0                          # a transpose: exact
    assert _rel(ds4[..., :Cn].permute(0, 3, 1, 2), a4.grad) < TOL                           # 1e-5
    assert float(ds3[..., Cn:].abs().max()) == 0.0 and float(ds4[..., Cn:].abs().max()) == 0.0


def _tf32(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,Cin,Cout,masked", [(4800, 256, 1024, True), (38400, 256, 1024, True), (1000, 64, 256, False),
                                               (130, 128, 512, True)])
def test_residual_epilogue_gemm_vs_torch(M, Cin, Cout, masked):
    """The `fast`-mode backward's fused "dx = dgrad(dy1) + [out > 0] * dout" kernel (it had no gate): TF32-exact
    operands, so the only difference from float64 is fp32 accumulation order."""
    from tinyfaces_b200 import ops
    g = torch.Generator().manual_seed(M + Cout)
    x = _tf32(torch.randn(1, 1, M, Cin, generator=g))
    w = _tf32(torch.randn(Cout, 1, Cin, generator=g) * 0.05)
    res = torch.randn(1, 1, M, Cout, generator=g)
    d = "cuda:0"
    mask = None
    gate = torch.ones(M, Cout, dtype=torch.bool)
    if masked:
        words = torch.randint(-2**31, 2**31 - 1, ((M * Cout + 31) // 32,), generator=g, dtype=torch.int64).to(torch.int32)
        mask = words.to(d)
        gate = _unpack(words, M * Cout).reshape(M, Cout)
    y = ops.conv2d_nhwc_res(x.to(d), w.to(d), res.to(d), mask)
    ref = x.reshape(M, Cin).double() @ w.reshape(Cout, Cin).double().t() + res.reshape(M, Cout).double() * gate.double()
    assert _rel(y.reshape(M, Cout), ref) < 2e-5                                            # fp32 accumulation of K <= 256 terms
    assert ops.gemm_error_flag() == 0
