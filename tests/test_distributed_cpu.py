"""CPU, world_size 2, gloo: the host-side logic of the two sharded paths (SUM gradient all-reduce of the
batch-sharded step; level-ordered candidate gather of the scale-sharded inference)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tinyfaces_b200.evaluation import gather_level_candidates
    from tinyfaces_b200.trainer import allreduce_gradients
    # --- SUM (not mean) gradient reduction
    ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    ps[0].grad = torch.full((5, 3), float(rank + 1))
    ps[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
    allreduce_gradients(ps)                    # ps[2] has no grad: skipped
    ok = torch.allclose(ps[0].grad, torch.full((5, 3), 3.0)) and torch.allclose(ps[1].grad, torch.arange(7.0) * 3) \
        and ps[2].grad is None
    # --- candidate gather keeps the `scales` (level) order on rank 0
    g = torch.Generator().manual_seed(0)
    levels = [(torch.rand(n, 4, generator=g, dtype=torch.float64), torch.rand(n, generator=g, dtype=torch.float64))
              for n in (5, 0, 11, 3, 7)]
    owner = [1, 0, 0, 1, 0]
    mine = {i: levels[i] for i in range(5) if owner[i] == rank}
    b, s = gather_level_candidates(mine, 5, dst=0)
    if rank == 0:
        ok = ok and torch.equal(b, torch.cat([l[0] for l in levels])) and torch.equal(s, torch.cat([l[1] for l in levels]))
    else:
        ok = ok and b is None
    # --- a rank that owns no level at all (more ranks than levels): it still takes part in both collectives
    mine = {0: levels[0], 2: levels[2]} if rank == 0 else {}
    b, s = gather_level_candidates(mine, 3, dst=0, device="cpu")
    if rank == 0:
        ok = ok and torch.equal(b, torch.cat([levels[0][0], levels[2][0]])) and s.shape[0] == 16
    else:
        ok = ok and b is None
    b, s = gather_level_candidates({}, 2, dst=0)                 # nobody has candidates
    if rank == 0:
        ok = ok and b.shape == (0, 4) and s.shape == (0,)
    with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(str(tmp_path), "ok%d" % r)).read() == "1"
