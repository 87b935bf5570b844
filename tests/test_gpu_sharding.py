"""reference-order sharding on real device work: two processes (gloo, both on cuda:0 — the exchange is host-side
metadata only) run their contiguous slices with sharding.run_chain_sharded; together they reproduce the
single-process chain bit for bit (images, labels, item order)."""
import os
import random

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

H, W, N = 200, 320, 7
KW = dict(num_rotations=2, include_original=True, pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)


def _inputs():
    from image_processor_pipeline_b200 import synth
    frames = np.stack([synth.frame(81, i, H, W) for i in range(N)])
    frames[3] = 0                                                  # one frame is keyed out entirely and disappears
    bgs = np.stack([synth.background(81, i, 240, 300) for i in range(3)])
    return frames, bgs


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from image_processor_pipeline_b200 import chain, sharding, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames, bgs = _inputs()
        lo, hi = sharding.shard_range(N, world, rank)
        params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **KW)
        res, own = sharding.run_chain_sharded(torch.from_numpy(frames[lo:hi]).to("cuda:0"), torch.from_numpy(bgs).to("cuda:0"),
                                              params, seed=4242)
        q.put((rank, own, res.images.cpu().numpy(), res.labels))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_the_single_process_chain():
    from image_processor_pipeline_b200 import chain, synth
    frames, bgs = _inputs()
    params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **KW)
    ref = chain.run_chain(torch.from_numpy(frames).to("cuda:0"), torch.from_numpy(bgs).to("cuda:0"), params, rng=random.Random(4242))
    ref_imgs, ref_labels = ref.images.cpu().numpy(), ref.labels
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = {}
    for _ in range(2):
        rank, own, imgs, labels = q.get(timeout=240)
        got[rank] = (own, imgs, labels)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    owned = got[0][0] + got[1][0]
    assert sorted(owned) == list(range(len(ref_labels))) and got[0][0] and got[1][0]
    for rank in (0, 1):
        own, imgs, labels = got[rank]
        for j, gi in enumerate(own):
            assert labels[j] == ref_labels[gi]
            assert np.array_equal(imgs[j], ref_imgs[gi]), (rank, gi)
