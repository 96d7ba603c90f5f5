"""CPU: host-side planning of the spatially tiled / sharded pyramid inference (evaluation.plan_bands / plan_jobs)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))


@pytest.mark.parametrize("H", [193, 500, 1250, 2500, 5000, 4999, 1237])
@pytest.mark.parametrize("nb", [1, 2, 3, 5, 8, 16])
def test_bands_cover_the_level_with_enough_halo(H, nb):
    from tinyfaces_b200.evaluation import HALO_PX, out_rows, plan_bands
    H3 = out_rows(H)
    bands = plan_bands(H, nb)
    assert bands[0][0] == 0 and bands[-1][1] == H3
    for (r0, r1, y0, y1), nxt in zip(bands, bands[1:] + [None]):
        assert r1 > r0 and (nxt is None or nxt[0] == r1)                  # disjoint, in order, complete
        assert r0 % 2 == 0 and y0 % 16 == 0 and 0 <= y0 < y1 <= H
        # every row of the band has its whole receptive field (radius 430 px) inside the tile or cut by the true border
        assert y0 == 0 or 8 * r0 - 430 >= y0
        assert y1 == H or 8 * (r1 - 1) + 430 < y1
        assert HALO_PX >= 430
        # the tile's own forward produces the rows the band needs
        assert out_rows(y1 - y0) >= r1 - y0 // 8
        if y1 == H:
            assert out_rows(y1 - y0) == H3 - y0 // 8


def test_job_plan_balances_the_five_level_pyramid():
    from tinyfaces_b200.evaluation import JOB_FIXED_MS, JOB_MS_PER_MPIX, plan_jobs
    shapes = [(312, 312), (625, 625), (1250, 1250), (2500, 2500), (5000, 5000)]

    def cost(h, w):
        return JOB_FIXED_MS + JOB_MS_PER_MPIX * h * w / 1e6          # the planner's estimate of a tile's forward time
    total = sum(cost(h, w) for h, w in shapes)
    for world in (1, 2, 4, 8):
        jobs = plan_jobs(shapes, world, spatial=True)
        assert [j[0] for j in jobs] == sorted(j[0] for j in jobs) and {j[0] for j in jobs} == set(range(5))
        load = [0.0] * world
        for lv, bi, r0, r1, y0, y1, owner in jobs:
            assert 0 <= owner < world
            load[owner] += cost(y1 - y0, shapes[lv][1])
        speedup = total / max(load)
        assert speedup >= {1: 1.0, 2: 1.6, 4: 2.6, 8: 4.1}[world] - 1e-9, (world, speedup)
        assert plan_jobs(shapes, world, spatial=True) is jobs          # cached
    # level sharding alone is capped by the 5000^2 level (23.5 of 31 estimated ms)
    jobs = plan_jobs(shapes, 8, spatial=False)
    assert len(jobs) == 5


def test_write_results_accepts_scores(tmp_path):
    """evaluation.py:89-114 with the score column restored (the shipped get_detections drops it, so the shipped writer
    raises): same file format, [K,5] or [K,4] + scores."""
    import numpy as np
    from tinyfaces_b200.evaluation import write_results
    dets = np.array([[10.4, 20.6, 30.2, 50.9], [0.0, 1.0, 2.0, 3.0]])
    f = write_results(dets, "0--Parade/0_Parade_x.jpg", "val", results_dir=str(tmp_path), scores=[0.5, 2.25])
    lines = open(f).read().splitlines()
    assert lines == ["0_Parade_x.jpg", "2", "10 21 21 31 0.5", "0 1 3 3 2.25"]
    f2 = write_results(np.concatenate([dets, [[0.5], [2.25]]], axis=1), "a/b.jpg", "val", results_dir=str(tmp_path))
    assert open(f2).read().splitlines()[2:] == lines[2:]
    with pytest.raises(ValueError):
        write_results(dets, "a/c.jpg", "val", results_dir=str(tmp_path))
