"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (dev container only).

    PYTHONPATH=/root/reference python -m oracle.make_golden

Imports /root/reference/tinyfaces (read-only) + the installed torchvision, runs
them on seeded synthetic inputs, and stores inputs + outputs as small fixtures.
/root/reference does not exist on the GPU box: tests only read the fixtures.
Versions are recorded in every file's ``meta`` entry.
"""
import json
import os
import sys

import numpy as np
import torch
import torchvision

sys.path.insert(0, "/root/reference")
from tinyfaces.models.model import DetectionModel            # noqa: E402
from tinyfaces.models.loss import DetectionCriterion         # noqa: E402
from tinyfaces.models import utils as ref_utils              # noqa: E402
from tinyfaces import evaluation as ref_eval                 # noqa: E402
from torchvision import transforms                           # noqa: E402

from . import synth                                          # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
META = json.dumps(dict(torch=torch.__version__, torchvision=torchvision.__version__,
                       numpy=np.__version__, reference="varunagrawal/tiny-faces-pytorch@a07e5da"))


def ref_model(sd):
    m = DetectionModel(pretrained_weights=None, num_templates=25)
    m.load_state_dict(sd, strict=True)
    return m


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name), meta=np.array(META), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


def golden_model():
    torch.manual_seed(0)
    for tag, gamma in (("g100", 1.0), ("g025", 0.25)):
        sd = synth.synthetic_state_dict(seed=1, bn3_gamma=gamma, beta_jitter=0.1)
        x = torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(2))
        m = ref_model(sd)
        # training mode: forward, loss-free backward against a fixed cotangent
        m.train()
        xr = x.clone().requires_grad_(False)
        out = m(xr)
        cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
        (out * cot).sum().backward()
        grads = {k: p.grad.numpy() for k, p in m.named_parameters() if p.grad is not None}
        sel = ["model.conv1.weight", "model.layer1.0.conv2.weight", "model.layer2.0.downsample.0.weight",
               "model.layer3.5.conv2.weight", "model.layer3.22.conv3.weight", "model.layer3.22.bn3.weight",
               "model.layer3.22.bn3.bias", "score_res3.weight", "score_res3.bias", "score_res4.weight",
               "score_res4.bias", "model.bn1.weight"]
        rs = m.state_dict()
        save("model_train_%s.npz" % tag, x=x.numpy(), out=out.detach().numpy(), cot=cot.numpy(),
             run_mean_l3=rs["model.layer3.22.bn3.running_mean"].numpy(),
             run_var_l3=rs["model.layer3.22.bn3.running_var"].numpy(),
             run_mean_bn1=rs["model.bn1.running_mean"].numpy(),
             run_var_bn1=rs["model.bn1.running_var"].numpy(),
             gamma=np.float32(gamma),
             **{"grad:" + k: (grads[k] if grads[k].size < 100000 else grads[k][:8]) for k in sel})
        # eval mode with calibrated statistics (oracle-side procedure, deterministic)
        xc = torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(4))
        sdc = synth.calibrate_running_stats(sd, xc)
        m2 = ref_model(sdc)
        m2.eval()
        xe = torch.randn(1, 3, 100, 129, generator=torch.Generator().manual_seed(5))
        with torch.no_grad():
            oe = m2(xe)
        save("model_eval_%s.npz" % tag, x=xe.numpy(), out=oe.numpy(), gamma=np.float32(gamma))


def golden_loss():
    r = np.random.RandomState(7)
    B, T, H, W = 2, 25, 12, 16
    out = (2.0 * r.randn(B, 5 * T, H, W)).astype(np.float32)
    u = r.rand(B, T, H, W)
    cm = np.zeros((B, T, H, W), np.float32)
    cm[u < 0.6] = -1
    cm[u > 0.8] = 1
    rm = (0.7 * r.randn(B, 4 * T, H, W)).astype(np.float32)
    crit = DetectionCriterion(25)
    o = torch.tensor(out, requires_grad=True)
    cmt = torch.tensor(cm.copy())
    np.random.seed(11)
    loss = crit(o, cmt, torch.tensor(rm))
    loss.backward()
    # the reference's post-sampling label map is not exposed; recover it from the masks
    labels_mask_cls = (crit.masked_class_loss.detach().numpy() != 0)
    save("loss_case0.npz", output=out, class_map=cm, regression_map=rm, np_seed=np.int64(11),
         total=loss.detach().numpy(), cls_sum=crit.masked_class_loss.sum().detach().numpy(),
         reg_sum=crit.masked_reg_loss.sum().detach().numpy(), class_map_after=cmt.numpy(),
         grad=o.grad.numpy(), cls_active=labels_mask_cls,
         class_avg=np.float64(crit.class_average.average), reg_avg=np.float64(crit.reg_average.average))
    # case 1: few labels (no sampling triggered), reg_weight 2
    cm1 = np.zeros((B, T, H, W), np.float32)
    cm1[u < 0.01] = -1
    cm1[u > 0.995] = 1
    crit = DetectionCriterion(25, reg_weight=2)
    o = torch.tensor(out, requires_grad=True)
    cmt = torch.tensor(cm1.copy())
    np.random.seed(12)
    loss = crit(o, cmt, torch.tensor(rm))
    loss.backward()
    save("loss_case1.npz", output=out, class_map=cm1, regression_map=rm, np_seed=np.int64(12),
         total=loss.detach().numpy(), cls_sum=crit.masked_class_loss.sum().detach().numpy(),
         reg_sum=crit.masked_reg_loss.sum().detach().numpy(), class_map_after=cmt.numpy(),
         grad=o.grad.numpy(), reg_weight=np.float32(2))


def golden_decode():
    templates = synth.load_templates()
    r = np.random.RandomState(21)
    H, W, T = 14, 30, 25
    for i, scale in enumerate((0.5, 1, 2.0, 0.7071067811865476)):
        sc = (1.5 * r.randn(1, H, W, T)).astype(np.float32)
        reg = (0.3 * r.randn(1, H, W, 4 * T)).astype(np.float32)
        prob = (1 / (1 + np.exp(-sc))).astype(np.float32)
        pin = prob.copy()
        boxes, scores = ref_utils.get_bboxes(sc, reg, pin, templates, 0.6, synth.RF, scale)
        save("decode_case%d.npz" % i, score_cls=sc, score_reg=reg, prob_cls=prob, prob_after=pin,
             scale=np.float64(scale), thresh=np.float64(0.6), boxes=boxes, scores=scores)
    # known answer: index (y=0,x=4,c=0), zero regression, scale 1
    sc = np.full((1, 1, 30, T), -10, np.float32)
    sc[0, 0, 4, 0] = 3.0
    reg = np.zeros((1, 1, 30, 4 * T), np.float32)
    prob = (1 / (1 + np.exp(-sc))).astype(np.float32)
    boxes, scores = ref_utils.get_bboxes(sc, reg, prob.copy(), templates, 0.5, synth.RF, 1)
    save("decode_known.npz", score_cls=sc, score_reg=reg, prob_cls=prob, boxes=boxes, scores=scores)


def golden_nms():
    for i, (n, thr) in enumerate(((2000, 0.3), (5000, 0.5), (300, 0.0))):
        boxes, scores = synth.synthetic_boxes(n, seed=30 + i, extent=400.0, dup_frac=0.05)
        keep = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        save("nms_case%d.npz" % i, boxes=boxes, scores=scores, thr=np.float64(thr), keep=keep)
    # float32 flavour
    boxes, scores = synth.synthetic_boxes(1500, seed=40, extent=300.0, dup_frac=0.05)
    b32, s32 = boxes.astype(np.float32), scores.astype(np.float32)
    keep = torchvision.ops.nms(torch.from_numpy(b32), torch.from_numpy(s32), 0.3).numpy()
    save("nms_case_f32.npz", boxes=b32, scores=s32, thr=np.float64(0.3), keep=keep)
    # large edge case: NaN / +-0 / +-inf scores, heavy ties, a few NaN coordinates -- big enough for the sort-and-sweep path
    boxes, scores = synth.synthetic_boxes(6000, seed=41, extent=500.0, dup_frac=0.05)
    r = np.random.RandomState(42)
    scores = np.round(scores, 2)
    scores[r.randint(0, 6000, 300)] = np.nan
    scores[r.randint(0, 6000, 300)] = 0.0
    scores[r.randint(0, 6000, 300)] = -0.0
    scores[r.randint(0, 6000, 50)] = np.inf
    scores[r.randint(0, 6000, 50)] = -np.inf
    scores[r.randint(0, 6000, 20)] = -np.nan
    for c in range(4):
        boxes[r.randint(0, 6000, 15), c] = np.nan
    keep = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.3).numpy()
    save("nms_edge.npz", boxes=boxes, scores=scores, thr=np.float64(0.3), keep=keep)
    keep = torchvision.ops.nms(torch.from_numpy(boxes), torch.full((6000,), 0.25, dtype=torch.float64), 0.3).numpy()
    save("nms_allequal.npz", boxes=boxes, scores=np.full(6000, 0.25), thr=np.float64(0.3), keep=keep)
    # known-answer semantics
    ka = []
    def run(b, s, t):
        k = torchvision.ops.nms(torch.tensor(b, dtype=torch.float64), torch.tensor(s, dtype=torch.float64), t)
        ka.append(dict(boxes=b, scores=s, thr=t, keep=k.tolist()))
    run([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [0, 0, 10, 10]], [0.5, 0.9, 0.5, 0.9], 0.5)
    run([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], 0.5)
    run([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], 0.4999)
    run([[1, 1, 1, 1], [1, 1, 1, 1]], [0.3, 0.7], 0.3)
    run([[0, 0, 4, 4], [1, 1, 3, 3], [5, 5, 6, 6], [0, 0, 4, 4.0001]], [0.1, 0.2, 0.3, 0.1], 0.2)
    # score-order edge semantics: torch.sort puts NaN first (all NaNs tie), -0.0 == +0.0, ties stable
    six = [[0, 0, 10, 10], [20, 20, 30, 30], [40, 40, 50, 50], [60, 60, 70, 70], [80, 80, 90, 90], [100, 100, 110, 110]]
    nan = float("nan")
    run(six, [0.0, -0.0, 0.0, -0.0, 0.5, nan], 0.3)
    run(six, [-0.0, 0.0, nan, -1.0, -nan, float("inf")], 0.3)
    run(six, [0.7] * 6, 0.3)
    same = [[0, 0, 10, 10]] * 3
    run(same, [0.5, nan, 0.9], 0.3)
    run(same, [0.0, -0.0, -0.0], 0.3)
    run(same, [-0.0, 0.0, -0.0], 0.3)
    run(same + [[0, 0, 10, 10.5]], [nan, nan, 1e300, -1e300], 0.3)
    run([[0, 0, 10, 10], [nan, 0, 10, 10], [0, 0, 10, 10], [0, 0, nan, 10]], [0.9, 0.8, 0.7, 0.6], 0.3)
    with open(os.path.join(OUT, "nms_known.json"), "w") as f:
        json.dump(dict(meta=json.loads(META), cases=ka), f, indent=1)
    print("wrote nms_known.json")


def golden_detections():
    templates = synth.load_templates()
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1)
    xc = torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(4))
    sd = synth.calibrate_running_stats(sd, xc)
    m = ref_model(sd)
    img = torch.rand(3, 200, 216, generator=torch.Generator().manual_seed(6))
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    for i, (thr, scales) in enumerate(((0.9, (0, 0.5, 1)),)):
        with torch.no_grad():
            dets = ref_eval.get_detections(m, img, templates, synth.RF, tf, prob_thresh=thr,
                                           nms_thresh=0.3, scales=scales, device=torch.device("cpu"))
        save("detections_case%d.npz" % i, img=img.numpy(), thresh=np.float64(thr),
             scales=np.array(scales, np.float64), dets=dets)


def _fingerprint(t):
    """Small witness of a seeded tensor that is regenerated (not stored) by the tests."""
    f = t.reshape(-1)
    return np.concatenate([f[:16].numpy().astype(np.float64), [float(f.double().sum()), float(f.double().abs().sum())]])


SEL_BASELINE = ["model.conv1.weight", "model.bn1.weight", "model.layer1.0.conv2.weight", "model.layer2.0.downsample.0.weight",
                "model.layer2.3.conv3.weight", "model.layer3.0.conv2.weight", "model.layer3.5.conv2.weight",
                "model.layer3.11.conv1.weight", "model.layer3.22.conv3.weight", "model.layer3.22.bn3.weight",
                "model.layer3.22.bn3.bias", "score_res3.weight", "score_res3.bias", "score_res4.weight", "score_res4.bias"]


def _sub(a, step):
    return np.ascontiguousarray(a[:, :, ::step, ::step])


def golden_baseline():
    """Goldens at the BASELINE.json shapes.  Inputs are regenerated from their seeds by the tests
    (a fingerprint is stored); outputs are stored as strided samples + the norms of the full map.
      cfg1: 1x3x500x500 -- eval forward (calibrated statistics) and train forward + gradients;
      cfg2: 1x3x960x1280 train forward + selected gradients; 8x3x960x1280 train forward only (the real GEMM sizes:
            tail split-K statistics, 2-CTA tiles; the CPU backward at batch 8 would need ~40 GB)."""
    sd = synth.synthetic_state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1)

    def train_case(name, B, H, W, seed, step, with_grads):
        x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed))
        m = ref_model(sd)
        m.train()
        kw = {}
        if with_grads:
            out = m(x)
            cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 1))
            (out * cot).sum().backward()
            for k, p in m.named_parameters():
                if k in SEL_BASELINE:
                    g = p.grad.numpy()
                    kw["grad:" + k] = g if g.size < 100000 else g[:8]
                    kw["gnorm:" + k] = np.float64(np.linalg.norm(g.astype(np.float64)))
            kw["cot_fp"] = _fingerprint(cot)
            out = out.detach()
        else:
            with torch.no_grad():
                out = m(x)
        rs = m.state_dict()
        o = out.numpy()
        save(name, x_fp=_fingerprint(x), shape=np.array([B, 3, H, W]), seed=np.int64(seed), step=np.int64(step),
             out_sub=_sub(o, step), out_l2=np.float64(np.linalg.norm(o.astype(np.float64))), out_max=np.float64(np.abs(o).max()),
             out_chan_sum=o.astype(np.float64).sum(axis=(0, 2, 3)),
             run_mean_l3=rs["model.layer3.22.bn3.running_mean"].numpy(), run_var_l3=rs["model.layer3.22.bn3.running_var"].numpy(),
             run_var_bn1=rs["model.bn1.running_var"].numpy(), **kw)

    train_case("cfg1_train.npz", 1, 500, 500, 50, 2, True)
    train_case("cfg2_b1_train.npz", 1, 960, 1280, 60, 4, True)
    train_case("cfg2_b8_fwd.npz", 8, 960, 1280, 70, 8, False)
    # cfg1 eval: BASELINE.json configs[0] (single 500x500 image, forward)
    xc = torch.randn(2, 3, 96, 136, generator=torch.Generator().manual_seed(4))
    sdc = synth.calibrate_running_stats(sd, xc)
    m = ref_model(sdc)
    m.eval()
    x = torch.randn(1, 3, 500, 500, generator=torch.Generator().manual_seed(80))
    with torch.no_grad():
        o = m(x).numpy()
    save("cfg1_eval.npz", x_fp=_fingerprint(x), shape=np.array([1, 3, 500, 500]), seed=np.int64(80), step=np.int64(2),
         out_sub=_sub(o, 2), out_l2=np.float64(np.linalg.norm(o.astype(np.float64))), out_max=np.float64(np.abs(o).max()),
         out_chan_sum=o.astype(np.float64).sum(axis=(0, 2, 3)))


def golden_targets():
    """Training targets from the reference's own DataProcessor.get_heatmaps / get_padding (processor.py:115-277).  The
    package __init__ of tinyfaces.datasets imports clustering code whose third-party dependencies (pyclust, pyclustering)
    are not installed; they are irrelevant to this path and stubbed out for the import."""
    from unittest import mock
    for name in ("pyclust", "pyclustering", "pyclustering.cluster", "pyclustering.cluster.kmedoids", "pyclustering.utils",
                 "pyclustering.utils.metric", "pyclustering.cluster.center_initializer"):
        sys.modules.setdefault(name, mock.MagicMock())
    from tinyfaces.datasets.processor import DataProcessor
    tpl = synth.load_templates()[:, :4]
    r = np.random.RandomState(5)
    cases = [
        (np.array([[100., 120., 180., 220.], [300., 50., 330., 90.], [10., 10., 400., 450.]]), [20, 30, 480, 470], 11),
        (np.concatenate([r.rand(9, 2) * 400, r.rand(9, 2) * 400], axis=1), [0, 0, 500, 500], 12),    # some boxes are invalid
        (np.zeros((0, 4)), [50, 60, 300, 310], 13),
        (np.array([[200., 200., 215., 218.], [200., 200., 215., 218.], [240., 100., 260., 130.]]), [0, 0, 500, 500], 14),   # duplicates
    ]
    b = cases[1][0]
    b[:, 2:] = b[:, :2] + (r.rand(9, 2) - 0.2) * 150
    for i, (boxes, paste, seed) in enumerate(cases):
        p = DataProcessor((500, 500), (63, 63), 0.7, 0.3, tpl, rf=synth.RF)
        pad = p.get_padding(paste)
        np.random.seed(seed)
        cls, reg, iou = p.get_heatmaps(boxes.copy(), pad)
        save("targets_case%d.npz" % i, bboxes=boxes, paste_box=np.array(paste), np_seed=np.int64(seed), pad_mask=pad,
             class_maps=cls.astype(np.int8), regress_maps=reg, iou_sum=np.float64(iou.sum()),
             iou_sample=iou.reshape(-1)[::97].copy(), iou_shape=np.array(iou.shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["model", "loss", "decode", "nms", "detections", "baseline", "targets"]
    for w in which:
        globals()["golden_" + w]()
