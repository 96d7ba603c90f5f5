"""GPU parity of the target-generation kernels (tf_heatmap_targets) against goldens produced by the reference's own
DataProcessor (oracle/make_golden.py targets) and against the CPU oracle on random cases: labels and the IoU volume
bit-exact (float64, reference operation order, identical np.random consumption), regression maps to 1e-13 (CUDA's log()
is within 1 ulp of the host's)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _processor(jitter="numpy"):
    from oracle import synth
    from tinyfaces_b200.targets import DataProcessor
    return DataProcessor((500, 500), (63, 63), 0.7, 0.3, synth.load_templates()[:, :4], rf=synth.RF, device="cuda:0", jitter=jitter)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_heatmaps_equal_reference_golden(case):
    g = np.load(os.path.join(G, "targets_case%d.npz" % case))
    p = _processor()
    pad = p.get_padding([int(v) for v in g["paste_box"]])
    assert np.array_equal(pad, g["pad_mask"])                                   # processor.py:115-155
    np.random.seed(int(g["np_seed"]))
    cls, reg, iou = p.get_heatmaps(g["bboxes"].copy(), pad)
    assert np.array_equal(cls.astype(np.int8), g["class_maps"])                 # labels: bit-exact
    assert tuple(iou.shape) == tuple(g["iou_shape"])
    assert np.array_equal(iou.reshape(-1)[::97], g["iou_sample"])               # IoU volume (+ noise): bit-exact
    assert abs(iou.sum() - float(g["iou_sum"])) <= 1e-9 * max(1.0, abs(float(g["iou_sum"])))
    scale = max(1.0, float(np.abs(g["regress_maps"]).max()))
    assert np.abs(reg - g["regress_maps"]).max() <= 1e-13 * scale               # tw / th go through log(): <= 1 ulp
    nt = 25
    assert np.array_equal(reg[:, :, : 2 * nt], g["regress_maps"][:, :, : 2 * nt])   # tx, ty: no transcendental -> exact
    # the host RNG was consumed exactly like the reference does (one rand(vsy, vsx, nt, ng) call, or none)
    next_draw = np.random.rand()
    np.random.seed(int(g["np_seed"]))
    ng = int(g["iou_shape"][3])
    if ng:
        np.random.rand(63, 63, 25, ng)
    assert next_draw == np.random.rand()


@pytest.mark.parametrize("seed,ng", [(1, 1), (2, 7), (3, 40)])
def test_heatmaps_equal_cpu_oracle_on_random_boxes(seed, ng):
    from oracle import synth, targets_oracle
    r = np.random.RandomState(seed)
    xy = r.rand(ng, 2) * 420
    wh = 8 + r.rand(ng, 2) * np.array([60, 300])[r.randint(0, 2, ng)][:, None]
    boxes = np.concatenate([xy, xy + wh], axis=1)
    p = _processor()
    pad = p.get_padding([int(r.randint(0, 60)), int(r.randint(0, 60)), 500 - int(r.randint(0, 60)), 500 - int(r.randint(0, 60))])
    np.random.seed(100 + seed)
    cls, reg, iou = p.get_heatmaps(boxes.copy(), pad)
    np.random.seed(100 + seed)
    ocls, oreg, oiou = targets_oracle.get_heatmaps(boxes.copy(), pad, synth.load_templates(), synth.RF, (63, 63), 0.7, 0.3)
    assert np.array_equal(cls, ocls) and np.array_equal(iou, oiou)
    assert np.abs(reg - oreg).max() <= 1e-13 * max(1.0, float(np.abs(oreg).max()))
    assert (cls == 1).sum() == (ocls == 1).sum() and (cls == 0).sum() == (ocls == 0).sum()


def test_device_layout_and_device_noise():
    """get_heatmaps_device: (C,H,W) float32 CUDA tensors = the transposed float copies of get_heatmaps; jitter='device'
    leaves np.random untouched and changes labels only where two ground-truth boxes tie."""
    g = np.load(os.path.join(G, "targets_case1.npz"))
    p = _processor()
    np.random.seed(int(g["np_seed"]))
    cm, rm = p.get_heatmaps_device(g["bboxes"].copy(), g["pad_mask"])
    assert cm.is_cuda and cm.dtype == torch.float32 and tuple(cm.shape) == (25, 63, 63) and tuple(rm.shape) == (100, 63, 63)
    assert np.array_equal(cm.cpu().numpy().astype(np.int8), g["class_maps"].transpose(2, 0, 1))
    assert np.abs(rm.cpu().numpy() - g["regress_maps"].transpose(2, 0, 1).astype(np.float32)).max() <= 1e-6
    pd = _processor(jitter="device")
    st = np.random.get_state()[1].copy()
    cls, reg, iou = pd.get_heatmaps(g["bboxes"].copy(), g["pad_mask"])
    assert np.array_equal(st, np.random.get_state()[1])
    assert (cls.astype(np.int8) != g["class_maps"]).mean() < 1e-3
