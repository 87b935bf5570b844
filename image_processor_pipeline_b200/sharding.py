"""Multi-GPU sharding of the chain: one process per GPU, contiguous slices, no
data-path collective.

Images are independent, so rank g of G owns frames [g*N/G, (g+1)*N/G).  Two modes:

* ``independent`` (bench, bulk production): every rank runs ``chain.run_chain`` on its
  slice with its own ``random.Random`` — nothing is exchanged at all.
* ``reference-order``: results identical to a single-process run whatever G is.  The
  reference draws its random parameters from ONE sequential stream whose consumption
  depends on sprite sizes, so the only thing ranks exchange is the per-image METADATA
  the host planner needs (stage-A component boxes, stage-B1 rotated boxes: 32 + 16 bytes
  per image, ``all_gather_object`` on the host side — gloo or NCCL).  Every rank then
  runs the same deterministic planner over ALL items and executes the items whose frame
  it owns.  No pixel ever crosses a GPU boundary.
"""
from __future__ import annotations

import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n work items for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of world {world}")
    return (rank * n) // world, ((rank + 1) * n) // world


def gather_rows(local: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-rank metadata rows in rank order (== input order for contiguous
    shards).  Works on any torch.distributed backend; a no-op without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    parts: List[Optional[np.ndarray]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    parts = [p for p in parts if p is not None and len(p)]
    return np.concatenate(parts, axis=0) if parts else local[:0]


def plan_global(info_all: np.ndarray, bbox_of_rot, params, seed_or_rng, n_bgs: int, bg_w: int, bg_h: int):
    """The deterministic, rank-independent host plan: rotation draws for every frame, then
    (once ``bbox_of_rot`` is known) every later draw.  Returns (rots, finals) exactly as
    ``chain.plan_rotations`` / ``chain.plan_items`` return them; each final item carries the
    GLOBAL frame index, so a rank keeps ``[f for f in finals if lo <= f['frame'] < hi]``."""
    from . import chain
    rng = _random.Random(seed_or_rng) if not hasattr(seed_or_rng, "uniform") else seed_or_rng
    rots = chain.plan_rotations(info_all, params, rng)
    if bbox_of_rot is None:
        return rots, None
    bbox = bbox_of_rot(rots) if callable(bbox_of_rot) else bbox_of_rot
    finals = chain.plan_items(rots, bbox, params, rng, n_bgs, bg_w, bg_h)
    return rots, finals


def run_chain_sharded(frames_local, backgrounds, params, seed: int, n_total: Optional[int] = None, group=None):
    """reference-order mode.  ``frames_local``: this rank's contiguous slice (CUDA uint8
    [n_local, H, W, 3]); ``backgrounds``: the full background set, replicated per GPU
    (CUDA uint8 [M, BH, BW, 3]).  Returns (ChainResult for the items this rank owns,
    their global item indices)."""
    import torch.distributed as dist
    from . import chain
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_local = frames_local.shape[0]
    counts = gather_rows(np.array([[n_local]], dtype=np.int64), group).ravel()
    lo = int(counts[:rank].sum())
    hi = lo + n_local
    return chain.run_chain(frames_local, backgrounds, params, rng=_random.Random(seed),
                           _shard=dict(lo=lo, hi=hi, gather=lambda a: gather_rows(a, group), world=world))
