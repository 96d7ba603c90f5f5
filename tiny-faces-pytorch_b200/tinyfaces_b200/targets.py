"""DataProcessor -- drop-in for the target-generation half of /root/reference/tinyfaces/datasets/processor.py:
``get_padding`` (:115-155) and ``get_heatmaps`` (:223-277, with ``get_regression`` :157-221 and
``compute_dense_overlap``, dense_overlap.py:4-75) run on the GPU (``tf_heatmap_targets``).  ``crop_image`` (image I/O
and augmentation) is not on this path and stays with the caller.

``get_heatmaps(bboxes, pad_mask)`` keeps the reference's numpy-in / numpy-out contract, including its host RNG
protocol: with ``jitter="numpy"`` (default) the tie-breaking noise is drawn by ``np.random.rand(vsy, vsx, nt, ng)``
exactly where the reference draws it (processor.py:203), so class maps are bit-identical under the same seed;
``jitter="device"`` draws the noise on the GPU instead (no host RNG, statistically equivalent).
``get_heatmaps_device`` returns CUDA tensors already in the (C, H, W) float32 layout the training step consumes
(wider_face.py:186-190 + trainer.py:74-76), skipping the device->host->device round trip altogether.
"""
import ctypes

import numpy as np
import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr


class DataProcessor:
    def __init__(self, input_size, heatmap_size, pos_thresh, neg_thresh, templates, img_means=None, rf=None,
                 device=None, jitter="numpy", seed=0):
        if jitter not in ("numpy", "device"):
            raise ValueError("jitter must be 'numpy' (reference RNG protocol) or 'device'")
        self.input_size = input_size
        self.heatmap_size = heatmap_size
        self.pos_thresh = pos_thresh
        self.neg_thresh = neg_thresh
        self.templates = np.asarray(templates, dtype=np.float64)
        self.rf = rf
        self.ofy, self.ofx = rf['offset']
        self.sty, self.stx = rf['stride']
        self.img_means = img_means or [0.485, 0.456, 0.406]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.jitter = jitter
        self._seed, self._calls = seed, 0

    def get_padding(self, paste_box):
        """processor.py:115-155 (pure index arithmetic on a 63x63x25 grid: stays in numpy)."""
        vsy, vsx = self.heatmap_size
        coarse_x, coarse_y = np.meshgrid(self.ofx + np.arange(vsx) * self.stx, self.ofy + np.arange(vsy) * self.sty)
        t = self.templates
        xx1 = coarse_x[:, :, None] + t[None, None, :, 0]
        yy1 = coarse_y[:, :, None] + t[None, None, :, 1]
        xx2 = coarse_x[:, :, None] + t[None, None, :, 2]
        yy2 = coarse_y[:, :, None] + t[None, None, :, 3]
        return (xx1 < paste_box[0] + 1) | (yy1 < paste_box[1] + 1) | (xx2 > paste_box[2]) | (yy2 > paste_box[3])

    def _run(self, bboxes, pad_mask, want_iou):
        vsy, vsx = self.heatmap_size
        nt = self.templates.shape[0]
        bboxes = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
        invalid = np.logical_or(bboxes[:, 2] <= bboxes[:, 0], bboxes[:, 3] <= bboxes[:, 1])         # processor.py:236-240
        bboxes = np.ascontiguousarray(np.delete(bboxes, np.where(invalid), axis=0))
        ng = bboxes.shape[0]
        dev = self.device
        jit = None
        if ng > 0 and self.jitter == "numpy":
            jit = torch.from_numpy(np.random.rand(vsy, vsx, nt, ng)).to(dev)                        # processor.py:203
        self._calls += 1
        b_d = torch.from_numpy(bboxes).to(dev) if ng else None
        pm = torch.from_numpy(np.ascontiguousarray(pad_mask).astype(np.uint8)).to(dev) if pad_mask is not None else None
        cls = torch.empty((vsy, vsx, nt), dtype=torch.float64, device=dev)
        reg = torch.empty((vsy, vsx, 4 * nt), dtype=torch.float64, device=dev)
        iou = torch.empty((vsy, vsx, nt, ng), dtype=torch.float64, device=dev) if want_iou else None
        sz = ctypes.c_size_t()
        check(lib().tf_targets_workspace_bytes(vsy, vsx, nt, ng, ctypes.byref(sz)), "tf_targets_workspace_bytes")
        ws = ops._workspace(dev, sz.value)
        tpl = np.ascontiguousarray(self.templates[:, :4])
        with torch.cuda.device(dev):
            check(lib().tf_heatmap_targets(ptr(b_d), ng, tpl.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nt, vsy, vsx,
                                           int(self.ofy), int(self.ofx), int(self.sty), int(self.stx), float(self.pos_thresh),
                                           float(self.neg_thresh), ptr(jit), (self._seed << 32) + self._calls, ptr(pm), ptr(cls),
                                           ptr(reg), ptr(iou) if (want_iou and ng) else None, ptr(ws), ws.numel(),
                                           stream_ptr(dev)), "tf_heatmap_targets")
        return cls, reg, iou

    def get_heatmaps(self, bboxes, pad_mask):
        """processor.py:223-277: (class_maps [vsy,vsx,nt], regress_maps [vsy,vsx,4nt], iou [vsy,vsx,nt,ng]) float64 ndarrays."""
        cls, reg, iou = self._run(bboxes, pad_mask, True)
        return cls.cpu().numpy(), reg.cpu().numpy(), iou.cpu().numpy()

    def get_heatmaps_device(self, bboxes, pad_mask):
        """CUDA float32 (class_map [nt,vsy,vsx], regression_map [4nt,vsy,vsx]) -- the layout / dtype of the training step."""
        cls, reg, _ = self._run(bboxes, pad_mask, False)
        return cls.permute(2, 0, 1).float().contiguous(), reg.permute(2, 0, 1).float().contiguous()
