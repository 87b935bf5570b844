"""Size-independent properties at BASELINE.json's full image sizes (the oracle would take minutes there):
involutions, idempotence, conservation checksums, identities with degenerate parameters.  Everything goes
through the C ABI (ops.py / chain.py); inputs are generated on the device."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ops = pytest.importorskip("image_processor_pipeline_b200.ops")
from image_processor_pipeline_b200 import chain, synth  # noqa: E402

DEV = "cuda:0"


def _noise(shape, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randint(0, 256, shape, dtype=torch.uint8, device=DEV, generator=g)


def test_c1_symmetries_1024_involution_and_checksum():
    """Config 1 (1024x1024 RGB): h∘h = v∘v = id, hv = h∘v = v∘h, and the byte checksum of every variant
    equals the input's (a permutation moves bytes, it does not change them) — over a batch of images."""
    for i in range(24):
        img = _noise((1024, 1024, 3), 100 + i)
        o, h, v, hv = ops.symmetries(img, ["o", "h", "v", "hv"])
        assert torch.equal(o, img)
        assert torch.equal(ops.flip_crop(h, flip_h=True), img) and torch.equal(ops.flip_crop(v, flip_v=True), img)
        assert torch.equal(ops.flip_crop(h, flip_v=True), hv) and torch.equal(ops.flip_crop(v, flip_h=True), hv)
        s = int(img.sum(dtype=torch.int64))
        assert all(int(t.sum(dtype=torch.int64)) == s for t in (h, v, hv))
        # per-row / per-column multisets survive the matching flip
        assert torch.equal(h.sum(dim=1, dtype=torch.int64), img.sum(dim=1, dtype=torch.int64))
        assert torch.equal(v.sum(dim=0, dtype=torch.int64), img.sum(dim=0, dtype=torch.int64))


def test_c1_square_crop_of_square_image_is_identity_and_1080p_window():
    img = _noise((1024, 1024, 3), 7)
    assert torch.equal(ops.flip_crop(img, 0, 0, 1024, 1024), img)                    # crop_square on a square image
    big = _noise((1080, 1920, 3), 8)
    for x0 in (0, 1, 419, 840):
        assert torch.equal(ops.flip_crop(big, x0, 0, 1080, 1080), big[:, x0:x0 + 1080])


@pytest.mark.parametrize("angle", [90.0, 180.0, 270.0])
def test_c2_right_angle_rotations_compose_to_identity_1080p(angle):
    """Config 2 (1920x1080 RGBA): PIL's transpose fast paths are exact permutations: k quarter turns that add up
    to 360 degrees give the image back; checksum conserved at every step."""
    img = ops.to_rgba(_noise((1080, 1920, 3), 21))
    s = int(img.sum(dtype=torch.int64))
    t = img
    for _ in range({90.0: 4, 180.0: 2, 270.0: 4}[angle]):            # 4 x 270 = 3 full turns
        t = ops.rotate_expand(t, angle)
        assert int(t.sum(dtype=torch.int64)) == s
    assert torch.equal(t, img)


@pytest.mark.parametrize("angle", [17.3, 45.0, 123.456, 301.0])
def test_c2_nearest_rotation_only_moves_pixels_1080p(angle):
    """Arbitrary angle, nearest neighbour: every opaque output pixel is a copy of an input pixel (checked with a
    per-pixel identifier image), transparent pixels are exactly zero, the opaque count is the source area within
    0.2 %, and the reported bbox is tight."""
    H, W = 1080, 1920
    ident = torch.arange(H * W, device=DEV, dtype=torch.int32)
    rgb = torch.stack([(ident & 255), ((ident >> 8) & 255), ((ident >> 16) & 255)], -1).to(torch.uint8).reshape(H, W, 3)
    out, bbox = ops.rotate_expand(ops.to_rgba(rgb), angle, with_bbox=True)
    a = out[..., 3]
    assert set(torch.unique(a).tolist()) <= {0, 255}
    assert int((out[a == 0] != 0).sum()) == 0
    ids = (out[..., 0].int() | (out[..., 1].int() << 8) | (out[..., 2].int() << 16))[a == 255]
    assert int(ids.max()) < H * W
    assert abs(int((a == 255).sum()) - H * W) <= 0.002 * H * W
    ys, xs = torch.nonzero(a, as_tuple=True)
    assert bbox == (int(xs.min()), int(ys.min()), int(xs.max()) + 1, int(ys.max()) + 1)


def test_c3_overlay_degenerate_sprites_1080p():
    """Config 3 (1920x1080 backgrounds): a fully transparent sprite leaves the background untouched, a fully opaque
    one replaces exactly its rectangle, and the fused square crop equals cropping afterwards."""
    bg = _noise((1080, 1920, 3), 31)
    spr = _noise((300, 400, 4), 32)
    clear = spr.clone(); clear[..., 3] = 0
    solid = spr.clone(); solid[..., 3] = 255
    assert torch.equal(ops.paste(bg, clear, (700, 333)), bg)
    out = ops.paste(bg, solid, (700, 333))
    assert torch.equal(out[333:633, 700:1100], solid[..., :3])
    out[333:633, 700:1100] = bg[333:633, 700:1100]
    assert torch.equal(out, bg)
    full = ops.paste(bg, spr, (700, 333))
    for x0 in (0, 421, 840):
        assert torch.equal(ops.paste(bg, spr, (700, 333), crop=(x0, 0, 1080, 1080)), full[:, x0:x0 + 1080])
    # blend is monotone in alpha between background and sprite colour
    lo, hi = torch.minimum(bg[333:633, 700:1100], spr[..., :3]), torch.maximum(bg[333:633, 700:1100], spr[..., :3])
    reg = full[333:633, 700:1100]
    assert bool(((reg >= lo) & (reg <= hi)).all())


def test_c4_key_and_keep_largest_512_batch_properties():
    """Config 4 (512x512): keying gives alpha in {0, 255} and never touches BGR; keep-largest is idempotent, only
    clears alpha, keeps exactly `area` pixels inside the reported box — over a batch of 128 frames."""
    n = 128
    frames = torch.from_numpy(np.stack([synth.frame(3, i % 8, 512, 512) for i in range(8)])).to(DEV).repeat(n // 8, 1, 1, 1)
    frames = torch.roll(frames, shifts=(37, 91), dims=(1, 2))
    keyed = ops.color_masks_to_alpha(frames, synth.REFERENCE_FILTERS)
    assert torch.equal(keyed[..., :3], frames)
    assert set(torch.unique(keyed[..., 3]).tolist()) <= {0, 255}
    once, info = ops.keep_largest_component(keyed)
    twice, info2 = ops.keep_largest_component(once)
    assert torch.equal(once, twice) and np.array_equal(info[:, :5], info2[:, :5])
    assert torch.equal(once[..., :3], frames)
    assert bool((once[..., 3] <= keyed[..., 3]).all())
    kept = (once[..., 3] != 0)
    assert np.array_equal(kept.sum(dim=(1, 2)).cpu().numpy(), info[:, 4])
    for i in (0, 17, 127):
        x, y, w, h = (int(v) for v in info[i, :4])
        assert int(kept[i, y:y + h, x:x + w].sum()) == int(info[i, 4])
        assert bool(kept[i, y].any()) and bool(kept[i, y + h - 1].any()) and bool(kept[i, :, x].any()) and bool(kept[i, :, x + w - 1].any())


def test_c5_chain_batch_properties_1080p():
    """Full chain on a batch of 64 1080p frames: one image and one label per frame, labels inside the unit square,
    every output differs from its background crop exactly inside the labelled box (plus nothing outside), and a
    second run with the same seed is identical (no race anywhere in the pipeline)."""
    n = 64
    base = [synth.frame(5, k, 1080, 1920) for k in range(4)]
    frames = torch.from_numpy(np.stack([np.roll(base[i % 4], (11 * i, 37 * i), (0, 1)) for i in range(n)])).to(DEV)
    bgs = _noise((n, 1080, 1920, 3), 55)
    params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, num_rotations=1, include_original=False,
                               pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False)
    random.seed(123)
    a = chain.run_chain(frames, bgs, params)
    random.seed(123)
    b = chain.run_chain(frames, bgs, params)
    assert a.labels == b.labels and torch.equal(a.images, b.images)
    assert a.images.shape == (n, 1080, 1080, 3) and len(a.labels) == n
    for i, (lab, rec) in enumerate(zip(a.labels, a.records)):
        cls, cx, cy, w, h = lab.split()
        cx, cy, w, h = float(cx), float(cy), float(w), float(h)
        assert cls == "0" and 0 < w <= 1 and 0 < h <= 1 and w / 2 <= cx <= 1 - w / 2 + 1e-6 and h / 2 <= cy <= 1 - h / 2 + 1e-6
        x0, y0, cs = rec["sq"]
        crop = bgs[rec["bg"], y0:y0 + cs, x0:x0 + cs]
        diff = (a.images[i] != crop).any(dim=2)
        ys, xs = torch.nonzero(diff, as_tuple=True)
        bx0, by0 = round((cx - w / 2) * cs), round((cy - h / 2) * cs)
        bx1, by1 = round((cx + w / 2) * cs), round((cy + h / 2) * cs)
        assert len(xs) > 0 and int(xs.min()) >= bx0 - 1 and int(xs.max()) <= bx1 and int(ys.min()) >= by0 - 1 and int(ys.max()) <= by1
