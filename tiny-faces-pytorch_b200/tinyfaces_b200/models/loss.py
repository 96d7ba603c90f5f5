"""DetectionCriterion -- drop-in for /root/reference/tinyfaces/models/loss.py:24-97.

OHEM, the masked SoftMargin + SmoothL1 sums and their gradient run in the library's kernels
(``tf_detloss_ohem``, ``tf_detloss_fwd_bwd``: one fused forward+backward pass).  Balance sampling runs on the device
by default (``sampler='device'``: ``tf_detloss_sample_device``, statistically equivalent to the reference, no host
round trip); ``sampler='numpy'`` reproduces the reference's host-side procedure draw for draw (identical
``np.random`` consumption, hence identical label maps under the same seed) for bit-exact comparisons.
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from .utils import balance_sampling


class AvgMeter:
    """loss.py:7-21 with the accumulation kept on the device (a running SUM updated in place, so that it also works
    inside a captured CUDA graph -- the replay driver only bumps ``num_averaged``); reading ``average`` is the only
    sync point."""

    def __init__(self):
        self.reset()

    def update(self, loss, size):
        if isinstance(loss, torch.Tensor):
            loss = loss.detach()
            if not isinstance(self._sum, torch.Tensor) or self._sum.device != loss.device:
                prev = float(self._sum) if self._sum is not None else 0.0
                self._sum = torch.full((), prev, dtype=torch.float64, device=loss.device)
            self._sum.add_(loss)
        else:
            self._sum = (self._sum if self._sum is not None else 0.0) + loss
        self.num_averaged += size

    @property
    def _average(self):
        """sum(loss_i) / sum(size_i): algebraically the reference's recurrence average = (n * average + loss) / (n + size),
        loss.py:13-17 (the batch loss is NOT weighted by its size there)"""
        if self._sum is None or self.num_averaged == 0:
            return 0
        return self._sum / self.num_averaged

    @property
    def average(self):
        a = self._average
        return float(a) if isinstance(a, torch.Tensor) else a

    def reset(self):
        if getattr(self, "_sum", None) is not None and isinstance(self._sum, torch.Tensor):
            self._sum.zero_()                    # keep the tensor: a captured graph holds its address
        else:
            self._sum = None
        self.num_averaged = 0


class _LossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, labels, regression_map, reg_weight):
        sums, grad = ops.detloss_fwd_bwd(output, labels, regression_map, reg_weight)
        ctx.save_for_backward(grad)
        sums32 = sums.to(torch.float32)
        total = sums32[0] + reg_weight * sums32[1]
        ctx.mark_non_differentiable(sums32)
        return total, sums32

    @staticmethod
    def backward(ctx, g_total, _g_sums):
        (grad,) = ctx.saved_tensors
        return grad * g_total, None, None, None


class DetectionCriterion(nn.Module):
    """The loss for the Tiny Faces detector (loss.py:24-97)."""

    def __init__(self, n_templates=25, reg_weight=1, pos_fraction=0.5, sampler="device", sample_size=256, seed=0):
        super().__init__()
        if sampler not in ("numpy", "device"):
            raise ValueError("sampler must be 'numpy' (reference RNG protocol) or 'device'")
        self.n_templates = n_templates
        self.reg_weight = reg_weight
        self.pos_fraction = pos_fraction
        self.sampler = sampler
        self.sample_size = sample_size
        self._seed = seed
        self._draws = None                  # device draw counter of the device sampler (a fresh sample per call / graph replay)
        self.class_average = AvgMeter()
        self.reg_average = AvgMeter()
        self.masked_class_loss = None       # 0-dim tensors: the masked SUMS (the reference's per-element maps are
        self.masked_reg_loss = None         # only ever consumed through .sum(), loss.py:87-91)
        self.total_loss = None

    def hard_negative_mining(self, output, class_map):
        """loss.py:59-63, in place on class_map.  Takes the full [B,5T,H,W] output (the kernel reads its
        classification channels in place instead of slicing a copy)."""
        ops.detloss_ohem_(output.detach().contiguous(), class_map)
        return class_map

    def balance_sample(self, class_map):
        """loss.py:47-57: per-image balance sampling; returns a NEW label tensor (the CUDA behaviour of the
        reference, where .cpu().numpy() is a copy)."""
        if self.sampler == "numpy":
            lab = class_map.cpu().numpy()
            for i in range(lab.shape[0]):
                balance_sampling(lab[i], pos_fraction=self.pos_fraction, sample_size=self.sample_size)
            return torch.from_numpy(lab).to(class_map.device)
        lab = class_map.clone()
        max_pos = int(self.sample_size * self.pos_fraction)
        max_neg = int(max_pos * (1 - self.pos_fraction) / self.pos_fraction)
        if self._draws is None or self._draws.device != lab.device:
            self._draws = torch.ones(1, dtype=torch.int64, device=lab.device)
        ops.detloss_sample_device_(lab, max_pos, max_neg, seed=self._seed << 32, counter=self._draws)
        return lab

    def forward(self, output, class_map, regression_map):
        if not output.is_cuda:
            raise RuntimeError("DetectionCriterion: output must be a CUDA tensor (no CPU fallback)")
        out_c = output.contiguous()
        cm = class_map if (class_map.is_contiguous() and class_map.dtype == torch.float32) else class_map.float().contiguous()
        ops.detloss_ohem_(out_c.detach(), cm)                       # loss.py:70 (in place)
        if cm is not class_map:
            class_map.copy_(cm)
        labels = self.balance_sample(cm)                            # loss.py:72
        total, sums = _LossFunction.apply(out_c, labels, regression_map.float().contiguous(), float(self.reg_weight))
        self.masked_class_loss = sums[0]
        self.masked_reg_loss = sums[1]
        self.total_loss = total
        self.class_average.update(sums[0], output.size(0))          # loss.py:90-91
        self.reg_average.update(sums[1], output.size(0))
        return total

    def loss_and_grad(self, output, class_map, regression_map):
        """forward() without autograd: returns (total loss, d total / d output).  The fused kernel produces both in one
        pass anyway; the autograd-free training step (trainer.train_step_flat, the CUDA-graph step) feeds the gradient
        straight into DetectionModel.backward_flat.  Same side effects as forward() (in-place OHEM, meters, attributes)."""
        if not output.is_cuda:
            raise RuntimeError("DetectionCriterion: output must be a CUDA tensor (no CPU fallback)")
        with torch.no_grad():
            out_c = output.detach().contiguous()
            cm = class_map if (class_map.is_contiguous() and class_map.dtype == torch.float32) else class_map.float().contiguous()
            ops.detloss_ohem_(out_c, cm)
            if cm is not class_map:
                class_map.copy_(cm)
            labels = self.balance_sample(cm)
            sums, grad = ops.detloss_fwd_bwd(out_c, labels, regression_map.float().contiguous(), float(self.reg_weight))
            sums32 = sums.to(torch.float32)
            total = sums32[0] + float(self.reg_weight) * sums32[1]
            self.masked_class_loss = sums32[0]
            self.masked_reg_loss = sums32[1]
            self.total_loss = total
            self.class_average.update(sums32[0], output.size(0))
            self.reg_average.update(sums32[1], output.size(0))
        return total, grad

    def reset(self):
        self.class_average.reset()
        self.reg_average.reset()
