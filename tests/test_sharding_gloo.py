"""N > 1 host logic on CPU: two gloo ranks gather the per-image metadata and each derives
the SAME global plan as a single process — results do not depend on the number of GPUs."""
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from image_processor_pipeline_b200 import chain, sharding
from test_host_logic import _fake_b1, _fake_stage_a

N = 11
PARAMS = dict(color_ranges_to_exclude_hsv=[(0, 0, 0, 180, 255, 150)], num_rotations=2, include_original=False,
              pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)


def _single():
    p = chain.ChainParams(**PARAMS)
    info = _fake_stage_a(N, 4)
    rng = random.Random(99)
    rots = chain.plan_rotations(info, p, rng)
    finals = chain.plan_items(rots, _fake_b1(rots), p, rng, 3, 1920, 1080)
    return [(f["frame"], f["rot_index"], f["sym"], f["bg"], f["new"], f["pos"], f["sq"], f["label"]) for f in finals]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = chain.ChainParams(**PARAMS)
        lo, hi = sharding.shard_range(N, world, rank)
        info_local = _fake_stage_a(N, 4)[lo:hi]                    # what stage A would return for the local frames
        info_all = sharding.gather_rows(info_local)
        rng = random.Random(99)
        rots = chain.plan_rotations(info_all, p, rng)
        local_rots = [r for r in rots if lo <= r["frame"] < hi]
        bbox_all = sharding.gather_rows(_fake_b1(local_rots))       # what stage B1 would return locally
        finals = chain.plan_items(rots, bbox_all, p, rng, 3, 1920, 1080)
        mine = [(f["frame"], f["rot_index"], f["sym"], f["bg"], f["new"], f["pos"], f["sq"], f["label"])
                for f in finals if lo <= f["frame"] < hi]
        q.put((rank, mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_plan_equals_single_process_plan():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=100) for _ in range(2))
    for pr in procs:
        pr.join(timeout=30)
        assert pr.exitcode == 0
    assert got[0] + got[1] == _single()
    assert got[0] and got[1]
