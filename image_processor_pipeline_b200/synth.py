"""Deterministic synthetic inputs for tests and benchmarks.

``frame(i)``: what a video frame of the reference's workflow looks like to the chain —
a dark backdrop that the first HSV filter of the reference's own demo list
(filtres_liste.py:186-190, V <= 150) keys out, one bright textured object (an ellipse of
random pose, colours mostly outside the other three filters, with a few yellow patches
that the other filters DO remove, i.e. holes) and ~1 % bright salt specks, so that the
key yields one large component plus O(10^2..10^4) small ones.
``background(i)``: uniform noise RGB.  Everything is uint8 and depends only on
(seed, i, shape).
"""
from __future__ import annotations

import numpy as np

# the only concrete HSV filter list in the reference (filtres_liste.py:186-190)
REFERENCE_FILTERS = [(0, 0, 0, 180, 255, 150), (15, 60, 200, 35, 255, 255),
                     (15, 30 * 2.55, 55 * 2.55, 30, 60 * 2.55, 80 * 2.55),
                     (15, 60 * 2.55, 60 * 2.55, 30, 75 * 2.55, 90 * 2.55)]


def noise(seed: int, i: int, shape) -> np.ndarray:
    """image i of a config: default_rng(seed*1_000_003 + i).integers(0, 256, shape)."""
    return np.random.default_rng(seed * 1_000_003 + i).integers(0, 256, shape, dtype=np.uint8)


def frame(seed: int, i: int, h: int, w: int, speck_density: float = 0.01) -> np.ndarray:
    """BGR frame (cv2 order) with a keyed-out backdrop, one object and salt specks."""
    rng = np.random.default_rng(seed * 1_000_003 + i)
    img = rng.integers(0, 120, (h, w, 3), dtype=np.uint8)                 # backdrop: V <= 150 -> keyed out
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    cy, cx = rng.uniform(0.35, 0.65) * h, rng.uniform(0.35, 0.65) * w
    ry, rx = rng.uniform(0.18, 0.30) * h, rng.uniform(0.12, 0.22) * w
    th = rng.uniform(0, np.pi)
    u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
    v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
    obj = (u / rx) ** 2 + (v / ry) ** 2 <= 1.0
    tex = rng.integers(0, 60, (h, w, 3), dtype=np.uint8)
    base = np.array([235, 170, 160], dtype=np.uint8)                       # bright, bluish in BGR: V > 150, H ~ 115
    img[obj] = (base - tex[obj] // 2).astype(np.uint8)
    # yellow patches inside the object that filters 2-4 remove (holes in the component)
    for _ in range(3):
        py, px = cy + rng.uniform(-0.5, 0.5) * ry, cx + rng.uniform(-0.5, 0.5) * rx
        pr = rng.uniform(0.02, 0.05) * min(h, w)
        patch = ((xx - px) ** 2 + (yy - py) ** 2 <= pr ** 2) & obj
        img[patch] = np.array([40, 215, 235], dtype=np.uint8)              # BGR yellow: H~27, S~212, V=235
    specks = rng.random((h, w)) < speck_density
    img[specks] = rng.integers(160, 256, (h, w, 3), dtype=np.uint8)[specks]   # bright specks survive the key
    return img


def background(seed: int, i: int, h: int, w: int) -> np.ndarray:
    return noise(seed + 7919, i, (h, w, 3))
