"""GPU parity tests of the individual C-ABI entry points against the oracle / golden fixtures."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _nms(boxes, scores, thr, algo=0):
    from tinyfaces_b200 import ops
    d = _dev()
    # algo: 0 auto, 1 blocked bit-matrix, 2 sort-and-sweep, 3 size-class grid (tf_nms_algo); no fallback here
    keep, count = ops.nms_device(torch.from_numpy(boxes).to(d), torch.from_numpy(scores).to(d), thr, algo)
    k = int(count.item())
    assert k >= 0, "sort-and-sweep edge list overflow"
    return keep[:k].cpu().numpy()


# ------------------------------------------------------------------------------------------- NMS
@pytest.mark.parametrize("algo", [1, 2, 3])
@pytest.mark.parametrize("name", ["nms_case0", "nms_case1", "nms_case2", "nms_case_f32", "nms_edge", "nms_allequal"])
def test_nms_golden_bit_exact(name, algo):
    """nms_edge / nms_allequal: NaN, +-0.0, +-inf scores, heavy ties, NaN coordinates (torch.sort order: NaN first, all NaNs
    tie, -0.0 == +0.0, stable) -- keep indices from torchvision.ops.nms itself."""
    g = np.load(os.path.join(G, name + ".npz"))
    assert np.array_equal(_nms(g["boxes"], g["scores"], float(g["thr"]), algo), g["keep"])


def test_nms_known_answers():
    with open(os.path.join(G, "nms_known.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        for algo in (1, 2, 3):
            k = _nms(np.array(c["boxes"], np.float64), np.array(c["scores"], np.float64), c["thr"], algo)
            assert k.tolist() == c["keep"], (c, algo)
    from tinyfaces_b200 import ops
    keep, count = ops.nms_device(torch.zeros((0, 4), dtype=torch.float64, device=_dev()),
                                 torch.zeros(0, dtype=torch.float64, device=_dev()), 0.3)
    assert keep.numel() == 0 and keep.dtype == torch.int64 and int(count.item()) == 0


@pytest.mark.parametrize("algo", [1, 2, 3])
@pytest.mark.parametrize("n,extent,thr", [(1, 10.0, 0.3), (63, 50.0, 0.3), (4097, 600.0, 0.3), (20000, 1500.0, 0.5),
                                          (40000, 1800.0, 0.3), (30000, 300.0, 0.3), (20000, 1200.0, 0.0)])
def test_nms_vs_c_oracle(n, extent, thr, algo):
    """bit-exact keep indices for both algorithms: the blocked bit-matrix path (incl. n > 32768) and the
    sort-and-sweep path (incl. a dense case with long suppression chains); ties and duplicates."""
    from oracle import nms_oracle, synth
    boxes, scores = synth.synthetic_boxes(n, seed=n, extent=extent, dup_frac=0.02)
    scores = np.round(scores, 3)            # many exact score ties
    assert np.array_equal(_nms(boxes, scores, thr, algo), nms_oracle.nms(boxes, scores, thr))


def test_nms_edge_semantics_float32_and_stream_order():
    """float32 flavour of the score-order edge cases against torchvision's CPU op semantics restated in numpy, and the
    stream-ordered contract: tf_nms only enqueues -- two calls on a side stream, results read after ONE synchronise."""
    from tinyfaces_b200 import ops
    g = np.load(os.path.join(G, "nms_edge.npz"))
    d = _dev()
    b = torch.from_numpy(g["boxes"]).to(d)
    s = torch.from_numpy(g["scores"]).to(d)
    side = torch.cuda.Stream(d)
    side.wait_stream(torch.cuda.current_stream(d))
    with torch.cuda.stream(side):
        k1, c1 = ops.nms_device(b, s, 0.3, 3)
        k2, c2 = ops.nms_device(b.clone(), torch.full_like(s, 0.25), 0.3, 2)       # same workspace, back to back: stream order
    side.synchronize()
    assert np.array_equal(k1[: int(c1.item())].cpu().numpy(), g["keep"])
    g2 = np.load(os.path.join(G, "nms_allequal.npz"))
    assert np.array_equal(k2[: int(c2.item())].cpu().numpy(), g2["keep"])
    st = ops.nms_sweep_stats(b.shape[0], 8, d)
    assert 0 < st["edges"] <= st["edge_capacity"] and st["rounds"] >= 1


def test_nms_sweep_overflow_is_flagged_and_falls_back():
    """A conflict list larger than the workspace's edge capacity: tf_nms_algo(2) reports -1 on the device, nms_keep re-runs
    the bit-matrix algorithm -- same answer as the oracle."""
    from oracle import nms_oracle
    from tinyfaces_b200 import ops
    n = 6000
    r = np.random.RandomState(3)
    c = r.rand(n, 2) * 6.0                                   # 6000 boxes of ~40 px in a 6 px square: ~n^2/2 conflicts
    wh = 40 + r.rand(n, 2)
    boxes = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
    scores = r.rand(n)
    d = _dev()
    b, s = torch.from_numpy(boxes).to(d), torch.from_numpy(scores).to(d)
    keep, count = ops.nms_device(b, s, 0.3, 2, exact_workspace=True)
    assert int(count.item()) == -1
    torch.cuda.synchronize()
    k = ops.nms_keep(b, s, 0.3)             # (falls back by itself if the shared workspace is small enough to overflow too)
    assert np.array_equal(k.cpu().numpy(), nms_oracle.nms(boxes, scores, 0.3))


@pytest.mark.parametrize("algo", [2, 3])
@pytest.mark.parametrize("seed,n", [(0, 20000), (1, 6000)])
def test_nms_multiscale_dense_vs_c_oracle(seed, n, algo):
    """Pyramid-like candidates: box sizes from 3 px to 900 px (eight size classes) packed into a 1250 px image, negative
    coordinates, a cluster of near-duplicates, a few degenerate / huge boxes -- the size-class grid and the 1-D sweep must
    both return the oracle's keep indices."""
    from oracle import nms_oracle
    r = np.random.RandomState(seed)
    size = np.exp(r.uniform(np.log(3.0), np.log(900.0), n))
    asp = np.exp(r.uniform(-0.7, 0.7, n))
    w, h = size * asp, size / asp
    c = r.rand(n, 2) * 1250.0 - 100.0
    boxes = np.stack([c[:, 0] - w / 2, c[:, 1] - h / 2, c[:, 0] + w / 2, c[:, 1] + h / 2], axis=1)
    k = n // 10
    boxes[:k] = boxes[k:2 * k] + r.randn(k, 4) * 0.5                      # near-duplicates: long suppression chains
    boxes[-5:, 2] = boxes[-5:, 0]                                         # zero width
    boxes[-8:-5, 2:] = boxes[-8:-5, :2] + 3000.0                          # larger than the image
    scores = np.round(r.rand(n), 3)
    got = _nms(boxes, scores, 0.3, algo)
    assert np.array_equal(got, nms_oracle.nms(boxes, scores, 0.3))


def test_nms_grid_flags_sizes_outside_its_classes():
    """A box wider than 2^16 (or narrower than 2^-16) has no size class: the grid variant flags the result (-1) and nms_keep
    falls back to the exact bit-matrix path."""
    from oracle import nms_oracle, synth
    from tinyfaces_b200 import ops
    boxes, scores = synth.synthetic_boxes(5000, seed=4, extent=400.0)
    boxes[7] = [0.0, 0.0, 2.0e5, 50.0]
    boxes[9] = [10.0, 10.0, 10.0 + 1e-6, 10.0 + 1e-6]
    d = _dev()
    b, s = torch.from_numpy(boxes).to(d), torch.from_numpy(scores).to(d)
    keep, count = ops.nms_device(b, s, 0.3, 3)
    assert int(count.item()) == -1
    torch.cuda.synchronize()
    assert np.array_equal(ops.nms_keep(b, s, 0.3).cpu().numpy(), nms_oracle.nms(boxes, scores, 0.3))


def test_nms_large_n_property():
    """N = 1e6 (beyond what the O(N^2) oracle can check): the kept set is conflict-free among the top kept boxes and
    every sampled removed box has a kept suppressor with a higher-or-equal score."""
    from oracle import synth
    n = 1000000
    boxes, scores = synth.synthetic_boxes(n, seed=1, dup_frac=0.001)
    keep = _nms(boxes, scores, 0.3)
    assert len(np.unique(keep)) == len(keep) and np.all(np.diff(scores[keep]) <= 0)          # unique, score-descending

    def iou(a, b):
        w = np.maximum(np.minimum(a[2], b[:, 2]) - np.maximum(a[0], b[:, 0]), 0)
        h = np.maximum(np.minimum(a[3], b[:, 3]) - np.maximum(a[1], b[:, 1]), 0)
        inter = w * h
        return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter)
    kb = boxes[keep]
    r = np.random.RandomState(0)
    for i in r.choice(len(keep), 200, replace=False):                  # kept boxes do not suppress each other
        ov = iou(kb[i], kb)
        ov[i] = 0
        assert not np.any((ov > 0.3) & (np.arange(len(keep)) < i))
    removed = np.setdiff1d(np.arange(n), keep)
    for j in r.choice(removed, 200, replace=False):                    # every removed box has a kept suppressor
        ov = iou(boxes[j], kb)
        assert np.any((ov > 0.3) & (scores[keep] >= scores[j]))


def test_nms_negative_threshold_and_ties_small():
    """thr < 0: every later box is suppressed (ovr = 0 > thr), as in the reference; handled by the bit-matrix path."""
    from oracle import nms_oracle, synth
    boxes, scores = synth.synthetic_boxes(700, seed=3, extent=200.0)
    for algo in (0, 1, 2, 3):
        k = _nms(boxes, scores, -0.1, algo)
        assert np.array_equal(k, nms_oracle.nms(boxes, scores, -0.1)) and len(k) == 1


def test_nms_degenerate_nothing_suppressed():
    from oracle import nms_oracle
    r = np.random.RandomState(5)
    boxes = np.repeat(r.rand(3000, 2) * 100, 2, axis=1)[:, [0, 2, 1, 3]]     # zero-area boxes: 0/0 -> NaN
    scores = r.rand(3000)
    k = _nms(boxes, scores, 0.3)
    assert np.array_equal(k, nms_oracle.nms(boxes, scores, 0.3)) and len(k) == 3000


# ------------------------------------------------------------------------------------------- decode
def _decode_nhwc(g, bug_compat=True):
    from oracle import decode_oracle, synth
    from tinyfaces_b200 import ops
    d = _dev()
    tpl = synth.load_templates()
    sc, reg, prob = (torch.from_numpy(g[k]).to(d) for k in ("score_cls", "score_reg", "prob_cls"))
    B, H, W, T = sc.shape
    inv = decode_oracle.invalid_ids(tpl, float(g["scale"]))
    mask = 0
    for i in inv:
        mask |= 1 << int(i)
    boxes, scores, src, count = ops.decode_device(sc, reg, prob, (H * W * T, W * T, T, 1), (H * W * 4 * T, W * 4 * T, 4 * T, 1),
                                                  B, H, W, T, tpl, float(g["thresh"]), mask if bug_compat else 0,
                                                  0 if bug_compat else mask, synth.RF, float(g["scale"]), want_src=True)
    n = int(count.item())
    return boxes[:n].cpu().numpy(), scores[:n].cpu().numpy(), src[:n].cpu().numpy()


def _assert_boxes_close(got, ref):
    """Box arithmetic is float64 in the reference's operation order; the only inexact step is numpy's
    *float32* exp (utils.py:87-88), which is not correctly rounded (up to ~2 ulp), so a coordinate may
    differ by a couple of float32 ulps of the box extent.  Tolerance: 4 * 2^-23 * extent."""
    ext = np.maximum(np.abs(ref[:, 2:] - ref[:, :2]), 1.0)
    tol = 4 * 2.0 ** -23 * np.concatenate([ext, ext], axis=1) + 1e-9
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) / tol))


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_decode_golden(case):
    g = np.load(os.path.join(G, "decode_case%d.npz" % case))
    boxes, scores, src = _decode_nhwc(g)
    assert boxes.shape == g["boxes"].shape                                  # same candidate count
    assert np.array_equal(scores.astype(np.float32), g["scores"][:, 0])    # same candidates, same (b,y,x,c) order
    flat = np.flatnonzero(g["prob_after"].reshape(-1) > np.float32(g["thresh"]))
    assert np.array_equal(src, flat)
    _assert_boxes_close(boxes, g["boxes"])


def test_decode_known_answer_and_nchw_sigmoid():
    from oracle import decode_oracle, synth
    from tinyfaces_b200 import ops
    d = _dev()
    tpl = synth.load_templates()
    r = np.random.RandomState(3)
    B, H, W, T = 1, 37, 53, 25
    out = (1.5 * r.randn(B, 5 * T, H, W)).astype(np.float32)
    out[:, T:] *= 0.2
    sc = np.ascontiguousarray(out[:, :T].transpose(0, 2, 3, 1))
    reg = np.ascontiguousarray(out[:, T:].transpose(0, 2, 3, 1))
    to = torch.from_numpy(out).to(d)
    prob = torch.sigmoid(to[:, :T]).permute(0, 2, 3, 1).contiguous().cpu().numpy()
    thr = 0.7
    # keep away from razor-edge probabilities so that sigmoid rounding cannot flip membership
    edge = np.abs(prob - np.float32(thr)) < 1e-5
    assert not edge.any()
    for scale in (0.5, 1.0, 2.0):
        rb, rs = decode_oracle.get_bboxes(sc, reg, prob.copy(), tpl, thr, synth.RF, scale)
        inv = decode_oracle.invalid_ids(tpl, scale)
        mask = sum(1 << int(i) for i in inv)
        hw = H * W
        boxes, scores, _, count = ops.decode_device(to, to[:, T:], None, (5 * T * hw, W, 1, hw), (5 * T * hw, W, 1, hw),
                                                    B, H, W, T, tpl, thr, mask, 0, synth.RF, scale)
        n = int(count.item())
        assert n == rb.shape[0]
        assert np.array_equal(scores[:n].cpu().numpy().astype(np.float32), rs[:, 0])
        _assert_boxes_close(boxes[:n].cpu().numpy(), rb)


# ------------------------------------------------------------------------------------------- loss
@pytest.mark.parametrize("case", [0, 1])
def test_loss_kernels_golden(case):
    from oracle import loss_oracle
    from tinyfaces_b200 import ops
    d = _dev()
    g = np.load(os.path.join(G, "loss_case%d.npz" % case))
    rw = float(g["reg_weight"]) if "reg_weight" in g.files else 1.0
    out = torch.from_numpy(g["output"]).to(d)
    cm = torch.from_numpy(g["class_map"].copy()).to(d)
    ops.detloss_ohem_(out, cm)
    cm_ref = g["class_map"].copy()
    loss_oracle.hard_negative_mining(g["output"][:, :25], cm_ref)
    assert np.array_equal(cm.cpu().numpy(), cm_ref)                          # OHEM, bit exact
    np.random.seed(int(g["np_seed"]))
    orc = loss_oracle.criterion(g["output"], g["class_map"].copy(), g["regression_map"], reg_weight=rw)
    labels = torch.from_numpy(orc["labels"]).to(d)
    sums, grad = ops.detloss_fwd_bwd(out, labels, torch.from_numpy(g["regression_map"]).to(d), rw)
    sums = sums.cpu().numpy()
    assert abs(sums[0] - g["cls_sum"]) <= 1e-5 * abs(g["cls_sum"])
    assert abs(sums[1] - g["reg_sum"]) <= 1e-5 * abs(g["reg_sum"])
    assert abs(sums[0] + rw * sums[1] - g["total"]) <= 1e-5 * abs(g["total"])
    np.testing.assert_allclose(grad.cpu().numpy(), g["grad"], rtol=1e-5, atol=1e-6)


def test_loss_device_sampler_counts():
    from tinyfaces_b200 import ops
    d = _dev()
    r = np.random.RandomState(0)
    B, T, H, W = 3, 25, 40, 50
    u = r.rand(B, T, H, W)
    lab = np.zeros((B, T, H, W), np.float32)
    lab[u < 0.9] = -1
    lab[u > 0.99] = 1
    lab[2][lab[2] > 0] = 0                       # image 2: no positives
    lab[1][(lab[1] < 0) & (r.rand(T, H, W) < 0.9995)] = 0   # image 1: few negatives
    t = torch.from_numpy(lab.copy()).to(d)
    ops.detloss_sample_device_(t, 128, 128, seed=1234)
    s = t.cpu().numpy()
    assert np.all((s == lab) | (s == 0))        # only ever zeroes labels
    for b in range(B):
        assert (s[b] > 0).sum() == min(128, (lab[b] > 0).sum())
        assert (s[b] < 0).sum() == min(128, (lab[b] < 0).sum())
    t2 = torch.from_numpy(lab.copy()).to(d)
    ops.detloss_sample_device_(t2, 128, 128, seed=99)
    assert not np.array_equal(t2.cpu().numpy(), s)           # a different seed picks a different subset


# ------------------------------------------------------------------------------------------- conv GEMMs
def _tf32(t):
    """round fp32 -> tf32 (10 explicit mantissa bits) so that tensor-core products are exact."""
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


CONV_CASES = [
    # B, H, W, Cin, Cout, k
    (1, 8, 16, 32, 64, 1),
    (2, 17, 23, 64, 128, 1),
    (1, 30, 40, 256, 256, 1),
    (2, 15, 20, 1024, 256, 1),
    (1, 13, 29, 256, 1024, 1),
    (1, 8, 16, 32, 64, 3),
    (2, 17, 23, 64, 64, 3),
    (1, 30, 40, 128, 128, 3),
    (2, 15, 20, 256, 256, 3),
    (1, 63, 63, 128, 128, 3),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", CONV_CASES)
def test_conv2d_nhwc_vs_torch_cpu(B, H, W, Cin, Cout, k):
    from tinyfaces_b200 import ops
    d = _dev()
    gen = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin + k)
    x = _tf32(torch.randn(B, Cin, H, W, generator=gen))
    w = _tf32(torch.randn(Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5)
    bias = torch.randn(Cout, generator=gen)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), padding=k // 2).float()
    xn = x.permute(0, 2, 3, 1).contiguous().to(d)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous().to(d)
    y = ops.conv2d_nhwc(xn, wp, k, bias=bias.to(d))
    torch.cuda.synchronize()
    assert ops.gemm_error_flag() == 0
    got = y.cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, "rel err %g" % err


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", CONV_CASES)
def test_conv2d_wgrad_vs_torch_cpu(B, H, W, Cin, Cout, k):
    from tinyfaces_b200 import ops
    d = _dev()
    gen = torch.Generator().manual_seed(7 + B * 1000 + H * 10 + Cin + k)
    x = _tf32(torch.randn(B, Cin, H, W, generator=gen))
    dy = _tf32(torch.randn(B, Cout, H, W, generator=gen))
    w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x.double(), w, padding=k // 2).backward(dy.double())
    ref = w.grad.float()
    xn = x.permute(0, 2, 3, 1).contiguous().to(d)
    dyn = dy.permute(0, 2, 3, 1).contiguous().to(d)
    dw = ops.conv2d_wgrad_nhwc(xn, dyn, k)
    torch.cuda.synchronize()
    assert ops.gemm_error_flag() == 0
    got = dw.cpu().reshape(Cout, k, k, Cin).permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, "rel err %g" % err


# shapes whose tile count does not fill a whole round over the 148 SMs and whose K loop is long enough to cut: the
# leftover tiles run as K slices that meet in the output through TMA reduce-add (tail-wave split-K)
SPLITK_CASES = [(2, 100, 100, 512, 256, 1), (1, 140, 150, 64, 128, 3), (3, 81, 80, 256, 256, 3), (1, 160, 125, 1024, 256, 1)]


@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k", SPLITK_CASES)
def test_conv2d_tail_split_k(B, H, W, Cin, Cout, k, accumulate):
    from tinyfaces_b200 import _lib, ops
    d = _dev()
    gen = torch.Generator().manual_seed(3 + B + H + Cin + k)
    x = _tf32(torch.randn(B, Cin, H, W, generator=gen))
    w = _tf32(torch.randn(Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5)
    ref = torch.nn.functional.conv2d(x, w, padding=k // 2)
    xn = x.permute(0, 2, 3, 1).contiguous().to(d)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous().to(d)
    y0 = torch.full((B, H, W, Cout), 0.5 if accumulate else 7.0, dtype=torch.float32, device=d)   # 7.0 must be overwritten
    outs = {}
    try:
        for split in (2, 0):                              # tf_debug_set(3, 2) disables the split: same numbers either way
            _lib.check(_lib.lib().tf_debug_set(3, split), "tf_debug_set")
            _lib.check(_lib.lib().tf_debug_set(1, accumulate), "tf_debug_set")
            outs[split] = ops.conv2d_nhwc(xn, wp, k, out=y0.clone()).cpu().permute(0, 3, 1, 2)
    finally:
        _lib.lib().tf_debug_set(3, 0)
        _lib.lib().tf_debug_set(1, 0)
    torch.cuda.synchronize()
    assert ops.gemm_error_flag() == 0
    if accumulate:
        ref = ref + 0.5
    scale = ref.abs().max().item()
    for split, got in outs.items():
        err = (got - ref).abs().max().item() / scale
        assert err < 3e-5, "split flag %d: rel err %g" % (split, err)        # fp32 CPU reference: 1e-5 noise of its own


STRIDED_CASES = [(1, 16, 16, 64, 64, 3), (2, 17, 23, 128, 128, 3), (1, 31, 40, 256, 256, 3), (2, 17, 23, 256, 512, 1),
                 (1, 30, 41, 512, 1024, 1)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", STRIDED_CASES)
def test_conv2d_stride2_fprop_and_wgrad_vs_torch_cpu(B, H, W, Cin, Cout, k):
    """stride-2 convolutions through the TMA traversal stride (no subsample / zero-insert passes)."""
    from tinyfaces_b200 import ops
    d = _dev()
    gen = torch.Generator().manual_seed(B + H + Cin + k)
    x = _tf32(torch.randn(B, Cin, H, W, generator=gen))
    w = _tf32(torch.randn(Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5).double().requires_grad_(True)
    ref = torch.nn.functional.conv2d(x.double(), w, stride=2, padding=k // 2)
    dy = _tf32(torch.randn(ref.shape, generator=gen))
    ref.backward(dy.double())
    xn = x.permute(0, 2, 3, 1).contiguous().to(d)
    wp = w.detach().float().permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous().to(d)
    y = ops.conv2d_nhwc_strided(xn, wp, k, 2)
    dw = ops.conv2d_wgrad_nhwc_strided(xn, dy.permute(0, 2, 3, 1).contiguous().to(d), k, 2)
    torch.cuda.synchronize()
    assert ops.gemm_error_flag() == 0
    got = y.cpu().permute(0, 3, 1, 2)
    assert tuple(got.shape) == tuple(ref.shape)
    e = (got - ref.detach().float()).abs().max().item() / ref.abs().max().item()
    gw = dw.cpu().reshape(Cout, k, k, Cin).permute(0, 3, 1, 2)
    ew = (gw - w.grad.float()).abs().max().item() / w.grad.abs().max().item()
    assert e < 2e-5 and ew < 2e-5, (e, ew)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(1, 16, 16, 64, 64, 3), (2, 17, 23, 128, 128, 3), (1, 31, 40, 256, 256, 3),
                                               (2, 18, 21, 256, 256, 3), (2, 17, 23, 256, 512, 1), (1, 30, 41, 512, 1024, 1)])
def test_conv2d_stride2_dgrad_parity_classes(B, H, W, Cin, Cout, k):
    """dx of a stride-2 conv from 4 parity-class GEMMs (no zero insertion) vs torch autograd; 1x1: accumulate form."""
    from tinyfaces_b200 import ops
    d = _dev()
    gen = torch.Generator().manual_seed(5 + B + H + Cin + k)
    x = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = _tf32(torch.randn(Cout, Cin, k, k, generator=gen) / (Cout * k * k) ** 0.5)
    y = torch.nn.functional.conv2d(x, w.double(), stride=2, padding=k // 2)
    dy = _tf32(torch.randn(y.shape, generator=gen))
    y.backward(dy.double())
    ref = x.grad.float()                                                     # [B, Cin, H, W]
    # dgrad packing: wp[ci][t][co] = w[co][ci][flipped tap t]
    wp = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, k * k, Cout).contiguous().to(d)
    dyn = dy.permute(0, 2, 3, 1).contiguous().to(d)
    if k == 3:
        dx = ops.conv2d_dgrad_s2_nhwc(dyn, wp, k, H, W, out=torch.full((B, H, W, Cin), 7.0, device=d))   # 7.0 must be overwritten
        base = 0.0
    else:
        dx = ops.conv2d_dgrad_s2_nhwc(dyn, wp, k, H, W, out=torch.full((B, H, W, Cin), 0.25, device=d), accumulate=True)
        base = 0.25
    torch.cuda.synchronize()
    assert ops.gemm_error_flag() == 0
    got = dx.cpu().permute(0, 3, 1, 2) - base
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 3e-5, "rel err %g" % err


def test_conv2d_3xtf32_parity_mode():
    """hi/lo split operands through the 3-segment K loop recover fp32-level accuracy on arbitrary fp32 data."""
    from tinyfaces_b200 import ops
    d = _dev()
    gen = torch.Generator().manual_seed(11)
    B, H, W, Cin, Cout, k = 2, 15, 20, 256, 256, 3
    x = torch.randn(B, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1).float()

    def split(t):
        hi = (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
        return hi, t - hi

    xn = x.permute(0, 2, 3, 1).contiguous()
    wp = w.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous()
    xh, xl = split(xn)
    wh, wl = split(wp)
    y1 = ops.conv2d_nhwc(xh.to(d), wh.to(d), k)
    y3 = ops.conv2d_nhwc(xh.to(d), wh.to(d), k, x_lo=xl.to(d), w_lo=wl.to(d))
    torch.cuda.synchronize()
    e1 = (y1.cpu().permute(0, 3, 1, 2) - ref).abs().max().item() / ref.abs().max().item()
    e3 = (y3.cpu().permute(0, 3, 1, 2) - ref).abs().max().item() / ref.abs().max().item()
    # the tensor core accumulates with truncation (~K/8 * 2^-24 relative), which bounds what the split can recover
    assert e3 < 5e-5 and e3 < e1 / 10, (e1, e3)
