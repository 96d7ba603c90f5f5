"""Run under torchrun on >= 2 GPUs:
 (1) sharded get_detections (by level, and by level + horizontal bands) equals the single-GPU result bit for bit;
 (2) data-parallel steps (bucketed SUM all-reduce overlapped with the backward, SGD per bucket), eager and replayed from
     a CUDA graph, leave identical parameters on every rank, and the reduced gradient equals the sum of the ranks' gradients."""
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
from tinyfaces_b200.optim import FlatSGD
from tinyfaces_b200.trainer import GraphedTrainStep, train_step_flat

STAGE = sys.argv[1] if len(sys.argv) > 1 else "all"          # inference | dp | graph | all
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])


def log(*a):
    print("[rank %d]" % rank, *a, flush=True)


torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import datetime
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
log("process group up")
if STAGE in ("inference", "all"):
  m = inference_bench.make_calibrated_model(dev, seed=0)
  tpl = inference_bench.load_templates()
  tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
  img = torch.rand(3, 1100, 1237, generator=torch.Generator().manual_seed(2))
  scales = (-1, 0, 0.5)
  thr = inference_bench.threshold_for(m, img, tf, scales, 30000, dev)
  with torch.no_grad():
      single = get_detections(m, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev) if rank == 0 else None
      for spatial, force in ((False, None), (True, None), (True, {2: 3, 1: 2})):
          sharded, jobs = get_detections_sharded(m, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev,
                                                 spatial=spatial, return_plan=True, force_bands=force)
          if rank == 0:
              same = np.array_equal(sharded, single)
              print("sharded inference (spatial=%s, forced bands %s): %d jobs, %d dets, identical to single-GPU: %s" % (spatial, force, len(jobs), len(sharded), same), flush=True)
              assert same
if STAGE == "inference":
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
# ---- data-parallel step: eager, then CUDA graph
torch.manual_seed(0)
model = DetectionModel(pretrained_weights=None, num_templates=25).to(dev).train()
crit = DetectionCriterion(25, sampler="device", seed=rank)
opt = FlatSGD(model, model.learnable_parameters(1e-6), momentum=0.9, weight_decay=5e-4, bucket_bytes=16 << 20)
x = synthetic.images(2, 128, 160, seed=10 + rank).to(dev)
cm, rm = synthetic.targets(2, 16, 20, seed=10 + rank, p_neg=0.7, p_pos=0.1)
cm, rm = cm.to(dev), rm.to(dev)
# the reduced gradient = sum over ranks of the local gradients (checked on one bucketed step with lr 0)
for g in opt.param_groups:
    g["lr_saved"], g["lr"] = g["lr"], 0.0
out = model.forward_train_flat(x)
_, grad = crit.loss_and_grad(out, cm.clone(), rm)
model.backward_flat(grad)
torch.cuda.synchronize()
local_grad = opt.flat.flat_grad.clone()
expect = local_grad.clone()
dist.all_reduce(expect, op=dist.ReduceOp.SUM)
crit2 = DetectionCriterion(25, sampler="device", seed=rank)
loss = train_step_flat(model, crit2, opt, x, cm.clone(), rm)
torch.cuda.synchronize()
err = float((opt.flat.flat_grad - expect).abs().max() / expect.abs().max())
for g in opt.param_groups:
    g["lr"] = g.pop("lr_saved")
opt._plans = None


def params_identical():
    flat = opt.flat.flat_param.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    ok = torch.tensor([1 if torch.equal(flat, ref) else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item())


log("bucketed step done, grad err %.2e" % err)
for _ in range(2):
    loss = train_step_flat(model, crit2, opt, x, cm.clone(), rm)
same_eager = params_identical()
log("eager steps identical:", same_eager)
graph_ok, same_graph = STAGE in ("graph", "all"), None
try:
    if not graph_ok:
        raise RuntimeError("graph stage skipped")
    gs = GraphedTrainStep(model, crit2, opt, x, cm, rm, warmup=1)
    for _ in range(3):
        loss = gs(x, cm, rm)
    torch.cuda.synchronize()
    same_graph = params_identical()
except Exception as ex:  # noqa: BLE001
    graph_ok = False
    print("rank %d: graph capture with NCCL failed: %s" % (rank, str(ex)[:300]), flush=True)
if rank == 0:
    print("data-parallel: reduced gradient vs sum of local gradients rel err %.2e (buckets %d); parameters identical on all %d ranks "
          "after eager steps: %s; after CUDA-graph steps: %s (capture ok: %s); loss %.4f"
          % (err, len(opt.flat.buckets), world, same_eager, same_graph, graph_ok, float(loss)), flush=True)
    # (two runs of the same step differ by ~5e-4 in the gradient: fp32 reduction order in the split-K / atomics paths moves
    #  activations by ~1e-7, which flips a few ReLU masks -- DESIGN.md section 2)
    assert err < 5e-3 and same_eager and (same_graph or not graph_ok)
try:
    del gs                                   # a live CUDA graph that captured NCCL kernels makes destroy_process_group() block
except NameError:
    pass
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
