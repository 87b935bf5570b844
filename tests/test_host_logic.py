"""Host-side scalar math and bookkeeping of the product against the oracle / Pillow. CPU only."""
import math
import random

import numpy as np
import pytest
from PIL import Image

from conftest import load_golden
from image_processor_pipeline_b200 import chain, hostmath, sharding
from oracle import intmodels as M, libcalls as L


@pytest.mark.parametrize("wh", [(64, 48), (1920, 1080), (333, 517), (1, 1), (2, 7)])
def test_rotation_plan_matches_pillow(wh):
    w, h = wh
    r = random.Random(w * 31 + h)
    for angle in [17.3, 45.0, 90.0, 180.0, 270.0, 0.0, 360.0, -30.0, 725.5] + [r.uniform(1, 359) for _ in range(20)]:
        plan = hostmath.rotation_plan(w, h, angle)
        kind, m, size = M.rotate_matrix(w, h, angle)
        assert plan.kind == kind and plan.size == size
        assert plan.size == Image.new("L", (w, h)).rotate(angle, expand=True).size
        if kind == "affine":
            assert plan.matrix == m and plan.fixed == M.affine_fixed_coeffs(m) and plan.fixed_ok == M.affine_fixed_ok(m, *size)


def test_overlay_size_and_crops_match_reference_expressions():
    r = random.Random(5)
    for _ in range(200):
        ow, oh = r.randint(1, 900), r.randint(1, 900)
        bw, bh = r.randint(50, 4000), r.randint(50, 4000)
        ratio = r.uniform(0.05, 1.2)
        assert hostmath.overlay_size(ow, oh, bw, bh, ratio) == L.overlay_size(ow, oh, bw, bh, ratio)
    assert hostmath.border_crop_rect(90, 60, (0.1, 5, 0.25, 7)) == (22, 6, 90 - 22 - 7, 60 - 6 - 5)
    with pytest.raises(ValueError):
        hostmath.border_crop_rect(90, 60, (-1, 0, 0, 0))
    with pytest.raises(ValueError):
        hostmath.border_crop_rect(90, 60, (30, 30, 0, 0))


def test_square_crop_bounds_match_oracle():
    r = np.random.default_rng(3)
    for _ in range(100):
        W, H = int(r.integers(40, 3000)), int(r.integers(40, 3000))
        n = int(r.integers(1, 5))
        boxes = np.column_stack([r.uniform(.2, .8, n), r.uniform(.2, .8, n), r.uniform(.02, .4, n), r.uniform(.02, .4, n)])
        try:
            _, cs, bx, by = L.square_crop_bounds(boxes, W, H)
        except RuntimeError:
            with pytest.raises(RuntimeError):
                hostmath.square_crop_bounds(boxes, W, H)
            continue
        assert hostmath.square_crop_bounds(boxes, W, H) == (cs, bx, by)


def test_hsv_range_packing_rounds_like_cv2():
    rs = hostmath.pack_hsv_ranges([(15, 30 * 2.55, 55 * 2.55, 30, 60 * 2.55, 80 * 2.55)], [(1, 2, 3, 4)])
    assert list(rs[0].lo) == [15, 76, 140] and list(rs[0].hi) == [30, 153, 204] and list(rs[0].zone) == [1, 2, 3, 4]
    rs = hostmath.pack_hsv_ranges([(300, 0, 50, 360, 100, 100)], None, use_gimp_scale=True)
    assert list(rs[0].lo) == [150, 0, 127] and list(rs[0].hi) == [180, 255, 255]      # 50*2.55 = 127.49999999999999 in float64
    with pytest.raises(ValueError):
        hostmath.pack_hsv_ranges([])
    with pytest.raises(ValueError):
        hostmath.pack_hsv_ranges([(0, 0, 0, 1, 1, 1)], [None, None])
    with pytest.raises(ValueError):
        hostmath.pack_hsv_ranges([(200, 0, 0, 10, 1, 1)])


def test_labels_known_answer_from_reference_self_test(tmp_path):
    from image_processor_pipeline_b200.transforms.labels import change_label_class
    g = load_golden("labels_0")
    (tmp_path / "in").mkdir(); (tmp_path / "out").mkdir()
    (tmp_path / "in" / "l.txt").write_text(g["meta"]["text_in"])
    res = change_label_class(tmp_path / "in" / "l.txt", [tmp_path / "out"], cls_mapping={0: 99, 1: 77})
    assert res.read_text() == g["meta"]["text_out"]
    assert change_label_class(tmp_path / "in" / "missing.txt", [tmp_path / "out"]) is None
    assert not (tmp_path / "out" / "missing.txt").exists()


def _fake_stage_a(n, seed):
    r = np.random.default_rng(seed)
    info = np.zeros((n, 8), np.int32)
    info[:, 0] = r.integers(0, 400, n); info[:, 1] = r.integers(0, 200, n)
    info[:, 2] = r.integers(100, 700, n); info[:, 3] = r.integers(100, 600, n)
    info[3 % n, 2:4] = 0                          # one frame keyed out entirely
    return info


def _fake_b1(rots):
    out = np.zeros((len(rots), 4), np.int32)
    for i, r in enumerate(rots):
        cw, ch = r["plan"].size
        out[i] = (1, 1, cw - 1, ch - 1) if cw > 2 and ch > 2 else (0, 0, cw, ch)
    return out


def test_plan_consumes_rng_like_the_reference_steps():
    """Draw order contract: N uniforms for the rotation step, then one
    sample per sprite, then (uniform, randint, randint) per variant, then 2 randint per item."""
    params = chain.ChainParams(color_ranges_to_exclude_hsv=[(0, 0, 0, 180, 255, 150)], num_rotations=2, include_original=True,
                               pool=["o", "h", "v", "hv"], choose_random=2)
    info = _fake_stage_a(5, 1)
    rng = random.Random(7)
    rots = chain.plan_rotations(info, params, rng)
    assert len(rots) == 4 * 3 and [r["rot_index"] for r in rots[:3]] == [0, 1, 2]
    chk = random.Random(7)
    angles = [chk.uniform(1.0, 359.0) for _ in range(8)]
    assert [r["angle"] for r in rots if r["angle"] is not None] == angles
    finals = chain.plan_items(rots, _fake_b1(rots), params, rng, 3, 1920, 1080)
    syms = []
    for _ in rots:
        f = chk.sample(["o", "h", "v", "hv"], 2)
        if "o" not in f:
            f.append("o")
        syms.append(sorted(f, key=chain.SYM_ORDER.get))
    assert [f["sym"] for f in finals] == [s for group in syms for s in group]
    assert [f["bg"] for f in finals] == [k % 3 for k in range(len(finals))]
    for f in finals:
        assert f["label"].endswith("\n") and len(f["label"].split()) == 5


def test_shard_ranges_partition_the_work_list():
    for n in (0, 1, 7, 1000, 10_000):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.shard_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_word_gather_addresses_stay_inside_aligned_extent():
    """The 3-byte-pixel gathers (chain.cu stage B2, batch.cu rotation) load the aligned 32-bit word holding a pixel's first
    byte and, when the pixel spills over (byte offset 2 or 3 inside the word), the next word.  Enumerated over every base
    alignment and pixel position: each loaded word contains at least one byte of the pixel, hence lies inside
    [floor4(buffer start), ceil4(buffer end)) — the access guarantee stated in include/ippgpu.h.  (compute-sanitizer is
    closed on the GPU pool of this build; this is the bounds check of our own.)"""
    for base in range(0, 8):
        for n_px in (1, 2, 5, 333 * 217):
            end = base + 3 * n_px
            lo_ok, hi_ok = base & ~3, (end + 3) & ~3
            for px in (set(range(min(n_px, 64))) | {n_px - 1, n_px - 2}) - {-1}:
                a = base + 3 * px
                words = [a & ~3]
                if (a & 3) * 8 > 8:
                    words.append((a & ~3) + 4)
                for w in words:
                    assert lo_ok <= w and w + 4 <= hi_ok
                    assert any(a <= b < a + 3 for b in range(w, w + 4)), "a loaded word without a byte of the pixel"
