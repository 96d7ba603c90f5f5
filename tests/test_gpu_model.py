"""GPU parity of DetectionModel forward / backward against the oracle and the reference-generated goldens.

Tolerances: the random-init net amplifies per-op rounding by 10x (bn3 gamma 0.25) to ~1000x
(gamma 1.0).  'parity' mode (3xTF32) must meet the north-star 1e-3 on the conditioned recipe; 'fast' mode
(1xTF32, what the reference itself runs on CUDA by default) is reported next to it.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _model(sd, precision):
    from tinyfaces_b200.models.model import DetectionModel
    m = DetectionModel(pretrained_weights=None, num_templates=25)
    m.load_state_dict(sd, strict=True)
    m.precision = precision
    return m.to("cuda:0")


def _record(name, rec):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "model_parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(test=name, **rec)) + "\n")


def test_state_dict_is_key_compatible():
    from oracle import synth
    sd = synth.synthetic_state_dict(seed=1)
    m = _model(sd, "fast")
    assert set(m.state_dict().keys()) == set(sd.keys())
    groups = m.learnable_parameters(1e-3)
    assert [g["lr"] for g in groups] == [1e-3, 1e-4, 1e-3, 0]


def test_layerwise_against_oracle_taps():
    """Every block output vs the oracle on the SAME weights/input: localises an error to a layer."""
    from oracle import model_oracle, synth
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1)
    x = torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(2))
    taps = {}
    with torch.no_grad():
        ref = model_oracle.forward(sd, x, training=True, taps=taps)
    m = _model(sd, "parity")
    m.train()
    with torch.no_grad():
        out = m(x.cuda())
    rec = {}
    names = [("pool", "stem")] + [("block%d.out" % i, k) for i, k in enumerate(
        ["model.layer1.%d" % j for j in range(3)] + ["model.layer2.%d" % j for j in range(4)] +
        ["model.layer3.%d" % j for j in range(23)])]
    worst = 0.0
    for mine, theirs in names:
        t = m.debug_tensor(mine).permute(0, 3, 1, 2).cpu().numpy()
        e = _rel(t, taps[theirs].numpy())
        rec[mine] = e
        worst = max(worst, e)
    rec["out"] = _rel(out.cpu().numpy(), ref.numpy())
    _record("layerwise_parity_g025", rec)
    assert rec["pool"] < 1e-4, rec
    assert worst < 1e-3 and rec["out"] < 1e-3, rec


@pytest.mark.parametrize("precision", ["parity", "mixed", "fast"])
def test_large_batch_shape_vs_oracle_autograd(precision):
    """2 x 400 x 400: large enough that the layer-1 GEMMs (157+ tiles on 148 SMs) take the tail split-K path, whose
    BN statistics come from a separate reduction -- forward, BN running stats and weight gradients vs the oracle."""
    from oracle import model_oracle, synth
    sd = synth.synthetic_state_dict(seed=3, bn3_gamma=0.25, beta_jitter=0.1)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 400, 400, generator=gen)
    keys = ["model.layer1.1.conv2.weight", "model.layer1.2.conv1.weight", "model.layer1.0.bn2.weight", "model.conv1.weight",
            "model.layer3.5.conv2.weight", "score_res3.weight"]
    state = dict(sd)
    for k in keys:
        state[k] = sd[k].clone().requires_grad_(True)
    new = {}
    ref = model_oracle.forward(state, x, training=True, new_stats=new)
    cot = torch.randn(ref.shape, generator=gen)
    (ref * cot).sum().backward()
    m = _model(sd, precision)
    m.train()
    out = m(x.cuda())
    (out * cot.cuda()).sum().backward()
    params = dict(m.named_parameters())
    rec = dict(precision=precision, out_max=_rel(out.detach().cpu().numpy(), ref.detach().numpy()),
               run_var=_rel(m.state_dict()["model.layer1.1.bn2.running_var"].cpu().numpy(), new["model.layer1.1.bn2.running_var"].numpy()))
    for k in keys:
        rec["grad:" + k] = _l2(params[k].grad.cpu().numpy(), state[k].grad.numpy())
    _record("large_shape", rec)
    if precision in ("parity", "mixed"):      # mixed: the same forward; its TF32 backward must stay inside the same gradient gate
        assert rec["out_max"] < 1e-3 and rec["run_var"] < 1e-4, rec
        assert max(v for k, v in rec.items() if k.startswith("grad:")) < 3e-2, rec
    else:
        assert rec["out_max"] < 3e-2 and rec["run_var"] < 1e-3, rec


@pytest.mark.parametrize("tag,gamma", [("g025", 0.25), ("g100", 1.0)])
@pytest.mark.parametrize("precision", ["parity", "mixed", "fast"])
def test_train_forward_backward_vs_reference_golden(tag, gamma, precision):
    from oracle import synth
    g = np.load(os.path.join(G, "model_train_%s.npz" % tag))
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=gamma, beta_jitter=0.1)
    m = _model(sd, precision)
    m.train()
    out = m(torch.from_numpy(g["x"]).cuda())
    (out * torch.from_numpy(g["cot"]).cuda()).sum().backward()
    rec = dict(tag=tag, precision=precision, out_max=_rel(out.detach().cpu().numpy(), g["out"]),
               out_l2=_l2(out.detach().cpu().numpy(), g["out"]))
    params = dict(m.named_parameters())
    for k in g.files:
        if k.startswith("grad:"):
            got = params[k[5:]].grad.cpu().numpy()
            ref = g[k]
            rec[k] = _l2(got[: ref.shape[0]], ref)
    sdm = m.state_dict()
    rec["run_mean_l3"] = _rel(sdm["model.layer3.22.bn3.running_mean"].cpu().numpy(), g["run_mean_l3"])
    rec["run_var_l3"] = _rel(sdm["model.layer3.22.bn3.running_var"].cpu().numpy(), g["run_var_l3"])
    rec["run_var_bn1"] = _rel(sdm["model.bn1.running_var"].cpu().numpy(), g["run_var_bn1"])
    assert int(sdm["model.bn1.num_batches_tracked"]) == 1
    _record("train_fwd_bwd", rec)
    assert params["score4_upsample.weight"].grad is None and params["model.fc.weight"].grad is None
    if precision in ("parity", "mixed") and tag == "g025":
        assert rec["out_max"] < 1e-3 and rec["out_l2"] < 1e-3, rec          # the north-star tolerance
        assert rec["run_var_bn1"] < 1e-4 and rec["run_mean_l3"] < 1e-3, rec
        assert max(v for k, v in rec.items() if k.startswith("grad:")) < 3e-2, rec
    elif precision == "fast" and tag == "g025":
        assert rec["out_max"] < 3e-2, rec
    else:                                     # gamma = 1.0: ~1000x error amplification, reported not gated tightly
        assert rec["out_max"] < (0.6 if precision == "fast" else 5e-2), rec


@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_eval_forward_vs_reference_golden(precision):
    from oracle import synth
    g = np.load(os.path.join(G, "model_eval_g025.npz"))
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1)
    xc = torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(4))
    sd = synth.calibrate_running_stats(sd, xc)
    m = _model(sd, precision)
    m.eval()
    with torch.no_grad():
        out = m(torch.from_numpy(g["x"]).cuda())
    assert tuple(out.shape) == g["out"].shape
    rec = dict(precision=precision, out_max=_rel(out.cpu().numpy(), g["out"]), out_l2=_l2(out.cpu().numpy(), g["out"]))
    _record("eval_fwd", rec)
    assert rec["out_max"] < (1e-3 if precision == "parity" else 3e-2), rec
