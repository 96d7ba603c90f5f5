"""CPU: pin every oracle function against fixtures produced by the reference
itself (oracle/make_golden.py).  The oracle is the checker for the GPU tests;
this file is what makes it trustworthy."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle, loss_oracle, model_oracle, nms_oracle, synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("tag,gamma", [("g100", 1.0), ("g025", 0.25)])
def test_model_oracle_train_forward_backward(tag, gamma):
    g = np.load(os.path.join(G, "model_train_%s.npz" % tag))
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=gamma, beta_jitter=0.1)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in sd.items()}
    new = {}
    out = model_oracle.forward(sd, torch.from_numpy(g["x"]), training=True, new_stats=new)
    # fp32 CPU convs: thread-count / blocking noise is amplified by the random-init net
    assert _rel(out.detach().numpy(), g["out"]) < 1e-3
    (out * torch.from_numpy(g["cot"])).sum().backward()
    for k in g.files:
        if k.startswith("grad:"):
            ref = g[k]
            got = sd[k[5:]].grad.numpy()[: ref.shape[0]] if ref.shape != tuple(sd[k[5:]].shape) else sd[k[5:]].grad.numpy()
            assert _rel(got, ref) < 2e-2, k
    assert _rel(new["model.layer3.22.bn3.running_mean"].numpy(), g["run_mean_l3"]) < 1e-3
    assert _rel(new["model.layer3.22.bn3.running_var"].numpy(), g["run_var_l3"]) < 1e-3
    assert _rel(new["model.bn1.running_var"].numpy(), g["run_var_bn1"]) < 1e-4


@pytest.mark.parametrize("tag,gamma", [("g100", 1.0), ("g025", 0.25)])
def test_model_oracle_eval(tag, gamma):
    g = np.load(os.path.join(G, "model_eval_%s.npz" % tag))
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=gamma, beta_jitter=0.1)
    xc = torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(4))
    sd = synth.calibrate_running_stats(sd, xc)
    with torch.no_grad():
        out = model_oracle.forward(sd, torch.from_numpy(g["x"]), training=False)
    assert out.shape == g["out"].shape
    assert _rel(out.numpy(), g["out"]) < 1e-3


@pytest.mark.parametrize("case", [0, 1])
def test_loss_oracle(case):
    g = np.load(os.path.join(G, "loss_case%d.npz" % case))
    cm = g["class_map"].copy()
    np.random.seed(int(g["np_seed"]))
    rw = float(g["reg_weight"]) if "reg_weight" in g.files else 1.0
    r = loss_oracle.criterion(g["output"], cm, g["regression_map"], reg_weight=rw, alias_cpu=True)
    assert np.array_equal(cm, g["class_map_after"])            # in-place OHEM + sampling (CPU view), bit exact
    assert abs(r["total"] - g["total"]) <= 1e-5 * abs(g["total"])
    assert abs(r["cls_sum"] - g["cls_sum"]) <= 1e-5 * abs(g["cls_sum"])
    assert abs(r["reg_sum"] - g["reg_sum"]) <= 1e-5 * abs(g["reg_sum"])
    assert np.allclose(r["grad"], g["grad"], rtol=1e-5, atol=1e-6)
    if "cls_active" in g.files:                                 # same sampled set <=> same RNG consumption
        act = (r["labels"] != 0) & (r["masked_cls"] != 0)
        assert np.array_equal(act, g["cls_active"])


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_decode_oracle(case):
    g = np.load(os.path.join(G, "decode_case%d.npz" % case))
    prob = g["prob_cls"].copy()
    boxes, scores = decode_oracle.get_bboxes(g["score_cls"], g["score_reg"], prob, synth.load_templates(),
                                             float(g["thresh"]), synth.RF, float(g["scale"]))
    assert np.array_equal(prob, g["prob_after"])                # the column-zeroing quirk, in place
    assert boxes.dtype == np.float64 and scores.dtype == np.float32
    assert np.array_equal(scores, g["scores"])
    assert np.array_equal(boxes, g["boxes"])                    # same numpy ops -> bit exact


def test_decode_known_answer():
    g = np.load(os.path.join(G, "decode_known.npz"))
    boxes, scores = decode_oracle.get_bboxes(g["score_cls"], g["score_reg"], g["prob_cls"].copy(),
                                             synth.load_templates(), 0.5, synth.RF, 1)
    assert np.array_equal(boxes, g["boxes"])
    assert np.allclose(boxes[0], [-55.83823529, -114.94411765, 117.83823529, 112.94411765])


@pytest.mark.parametrize("name", ["nms_case0", "nms_case1", "nms_case2", "nms_edge", "nms_allequal"])
def test_nms_oracle_c_and_numpy(name):
    g = np.load(os.path.join(G, name + ".npz"))
    k = nms_oracle.nms(g["boxes"], g["scores"], float(g["thr"]))
    assert np.array_equal(k, g["keep"])
    if len(g["scores"]) <= 2000:
        assert np.array_equal(nms_oracle.nms_numpy(g["boxes"], g["scores"], float(g["thr"])), g["keep"])


def test_nms_known_answers():
    with open(os.path.join(G, "nms_known.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        k = nms_oracle.nms(np.array(c["boxes"], np.float64), np.array(c["scores"], np.float64), c["thr"])
        assert k.tolist() == c["keep"], c
    assert nms_oracle.nms(np.zeros((0, 4)), np.zeros(0), 0.3).shape == (0,)


@pytest.mark.parametrize("name,training", [("cfg1_eval", False), ("cfg1_train", True)])
def test_model_oracle_at_cfg1_shape(name, training):
    """BASELINE configs[0] (1x3x500x500 forward on CPU): the oracle against the reference-generated golden samples."""
    import torch
    from oracle import model_oracle, synth
    g = np.load(os.path.join(G, name + ".npz"))
    B, C, H, W = (int(v) for v in g["shape"])
    x = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(int(g["seed"])))
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1)
    if not training:
        sd = synth.calibrate_running_stats(sd, torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(4)))
    with torch.no_grad():
        out = model_oracle.forward(sd, x, training=training).numpy()
    step = int(g["step"])
    assert np.abs(out[:, :, ::step, ::step] - g["out_sub"]).max() <= 1e-3 * float(g["out_max"])
    assert abs(np.linalg.norm(out.astype(np.float64)) - float(g["out_l2"])) <= 1e-3 * float(g["out_l2"])


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_targets_oracle_against_reference_golden(case):
    """get_heatmaps / get_regression / compute_dense_overlap (processor.py:157-277, dense_overlap.py:4-75): the vectorised
    oracle against the reference's own output, same np.random seed -> identical labels, IoU volume and regression maps."""
    from oracle import targets_oracle
    g = np.load(os.path.join(G, "targets_case%d.npz" % case))
    np.random.seed(int(g["np_seed"]))
    cls, reg, iou = targets_oracle.get_heatmaps(g["bboxes"].copy(), g["pad_mask"], synth.load_templates(), synth.RF, (63, 63), 0.7, 0.3)
    assert np.array_equal(cls.astype(np.int8), g["class_maps"])
    assert tuple(iou.shape) == tuple(g["iou_shape"]) and np.array_equal(iou.reshape(-1)[::97], g["iou_sample"])
    assert np.abs(reg - g["regress_maps"]).max() <= 1e-12
