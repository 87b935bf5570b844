"""Parity of the batched operators (per-item parameter arrays, one launch per batch) against the oracle:
the reference's cv2 / Pillow calls per image, with the random draws taken from the same seeded generator."""
import random

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import libcalls as L

pytestmark = pytest.mark.gpu

batch = pytest.importorskip("image_processor_pipeline_b200.batch")
ops = pytest.importorskip("image_processor_pipeline_b200.ops")
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("case", [((6, 40, 64, 3), 21, 17), ((3, 1080, 1920, 3), 1080, 1080), ((5, 33, 47, 4), 9, 33), ((4, 20, 20, 1), 20, 20)])
def test_crop_batch(case):
    shape, ow, oh = case
    r = np.random.default_rng(3)
    imgs = r.integers(0, 256, shape, dtype=np.uint8)
    n, H, W, _ = shape
    xy = np.stack([r.integers(0, W - ow + 1, n), r.integers(0, H - oh + 1, n)], axis=1)
    out = host(batch.crop_batch(dev(imgs), xy, ow, oh))
    for i in range(n):
        assert np.array_equal(out[i], imgs[i, xy[i, 1]:xy[i, 1] + oh, xy[i, 0]:xy[i, 0] + ow])
    with pytest.raises(ValueError):
        batch.crop_batch(dev(imgs), xy + W, ow, oh)


@pytest.mark.parametrize("shape", [(7, 61, 97, 3), (5, 50, 31, 4), (2, 1080, 1920, 3)])
def test_rotate_batch_nearest_bit_exact(shape):
    """rotations.py:55,89,96-99 per image: draws, canvases, getbbox and tight crops equal Pillow's."""
    r = np.random.default_rng(11)
    imgs = r.integers(0, 256, shape, dtype=np.uint8)
    if shape[3] == 4:
        imgs[..., 3] = np.where(r.random(shape[:3]) < 0.3, 0, imgs[..., 3])
    random.seed(5)
    res = batch.rotate_batch(dev(imgs), random, 1.0, 359.0)
    state = random.getstate()
    random.seed(5)
    for i in range(shape[0]):
        angle = random.uniform(1.0, 359.0)
        assert angle == res.angles[i]
        rgba = L.to_rgba(imgs[i])
        ref = L.rotate_expand(rgba, angle)
        assert np.array_equal(host(res.canvas(i)), ref)
        fit, bbox = L.rotate_and_fit(rgba, angle)
        assert np.array_equal(host(res.fit(i)), fit)
        if bbox is not None:
            assert tuple(int(v) for v in res.bbox[i]) == tuple(bbox)
    assert random.getstate() == state


def test_rotate_batch_special_angles_and_fixed_list():
    r = np.random.default_rng(12)
    imgs = r.integers(0, 256, (5, 37, 53, 4), dtype=np.uint8)
    angles = [0.0, 90.0, 180.0, 270.0, 33.25]
    res = batch.rotate_batch(dev(imgs), angles=angles)
    for i, a in enumerate(angles):
        assert np.array_equal(host(res.canvas(i)), L.rotate_expand(imgs[i], a))


@pytest.mark.parametrize("shape", [(4, 217, 333, 3), (3, 120, 90, 4), (1, 1080, 1920, 3)])
def test_rotate_batch_bilinear_within_one_lsb(shape):
    """Bilinear mode (not used by the reference; oracle = the same PIL call with resample=BILINEAR).
    Tolerance: <= 1 LSB per channel (RGB sources interpolate in single precision); RGBA sources use libImaging's
    double-precision interpolation and are expected equal."""
    r = np.random.default_rng(13)
    imgs = r.integers(0, 256, shape, dtype=np.uint8)
    angles = [float(a) for a in r.uniform(1, 359, shape[0])]
    res = batch.rotate_batch(dev(imgs), angles=angles, resample="bilinear")
    for i, a in enumerate(angles):
        ref = L.rotate_expand(L.to_rgba(imgs[i]), a, Image.Resampling.BILINEAR)
        got = host(res.canvas(i))
        diff = np.abs(got.astype(np.int16) - ref.astype(np.int16))
        assert diff.max() <= 1, diff.max()
        if shape[3] == 4:
            assert diff.max() == 0
        else:
            assert (diff > 0).mean() < 1e-3


@pytest.mark.parametrize("src", [(48, 64), (384, 512), (20, 31)])
def test_resize_lanczos_batch_bit_exact(src):
    sh, sw = src
    r = np.random.default_rng(21)
    sprites = r.integers(0, 256, (3, sh, sw, 4), dtype=np.uint8)
    sprites[1, ..., 3] = np.where(r.random((sh, sw)) < 0.5, 0, 255)
    sizes = [(sw // 2 + 3, sh // 2 + 1), (sw, sh // 3 + 1), (sw * 2 + 1, sh + 5), (sw, sh), (max(1, sw // 7), max(1, sh // 5)), (sw + 9, sh)]
    which = [0, 1, 2, 0, 1, 2]
    for device_tables in (False, True):
        outs = batch.resize_lanczos_batch(dev(sprites), sizes, which, device_tables=device_tables)
        for o, size, s in zip(outs, sizes, which):
            assert np.array_equal(host(o), L.resize_lanczos(sprites[s], size)), (size, device_tables)
    one = ops.resize_lanczos(dev(sprites[0]), sizes[0])                 # the single-image op runs the same tensor-core passes
    assert np.array_equal(host(one), L.resize_lanczos(sprites[0], sizes[0]))


@pytest.mark.parametrize("geom", [((96, 128), (270, 480), 9), ((384, 512), (1080, 1920), 5)])
def test_overlay_batch_equals_reference_calls(geom):
    """overlays.py:104-149 per (sprite, background) pair: same draws, same composite, same label text."""
    (sh, sw), (BH, BW), n = geom
    r = np.random.default_rng(31)
    sprites = r.integers(0, 256, (3, sh, sw, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:sh, 0:sw]
    ramp = np.clip(255 * (1.2 - np.hypot((xx - sw / 2) / (sw / 2), (yy - sh / 2) / (sh / 2))), 0, 255).astype(np.uint8)
    sprites[..., 3] = ramp                                               # radial alpha ramp with 0 / 255 plateaus
    bgs = r.integers(0, 256, (4, BH, BW, 3), dtype=np.uint8)
    random.seed(8)
    res = batch.overlay_batch(dev(sprites), dev(bgs), n, random, yolo_class_id=2)
    state = random.getstate()
    random.seed(8)
    assert len(res.labels) == n == res.images.shape[0]
    out = host(res.images)
    for k in range(n):
        comp, label, _ = L.overlay_onto_background(sprites[k % 3], bgs[k % 4], yolo_class_id=2)
        assert label == res.labels[k]
        assert np.array_equal(out[k], comp)
    assert random.getstate() == state


@pytest.mark.parametrize("shape", [(6, 96, 130), (4, 512, 512), (3, 61, 77)])
def test_key_and_keep_largest_equals_the_two_reference_steps(shape):
    """filtres_liste.py:90-134 then pixels_isolés.py:29-81 per image, fused on the device (frame -> bits -> CCL -> BGRA)."""
    from image_processor_pipeline_b200 import synth
    n, H, W = shape
    frames = np.stack([synth.frame(70, i, H, W, speck_density=0.02) for i in range(n)])
    frames[-1] = synth.frame(70, 99, H, W)
    frames[-1][frames[-1] > 0] //= 3                                   # everything dark: the key removes the whole image
    out, info = batch.key_and_keep_largest(dev(frames), synth.REFERENCE_FILTERS)
    out = host(out)
    for i in range(n):
        keyed = L.color_masks_to_alpha(frames[i], list(synth.REFERENCE_FILTERS))
        x, y, w, h = (int(v) for v in info[i, :4])
        if not (keyed[..., 3] > 1).any():
            assert w == 0 and h == 0 and not out[i, ..., 3].any()
            continue
        ref = L.keep_largest_component(keyed)
        assert np.array_equal(out[i, y:y + h, x:x + w], ref)
        assert np.array_equal(out[i, ..., :3], frames[i])


@pytest.mark.parametrize("shape", [(7, 48, 70), (5, 120, 90), (3, 1080, 1080), (4, 33, 5)])
def test_enhance_batch_bit_exact(shape):
    """tranfo.py:37-51 per image (Pillow ImageEnhance blends, GaussianBlur box passes, per-channel point tables): the batch on
    the device equals the reference's calls, with the draws taken from the same generator in the same order."""
    n, H, W = shape
    r = np.random.default_rng(51)
    imgs = r.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    imgs[0] = (imgs[0] // 16) + 100                                   # a low-contrast image: extrapolated blends get clipped
    random.seed(17)
    do_blur, do_rgb = batch.sample_flags(n, random)
    do_blur[0], do_rgb[min(1, n - 1)], do_blur[-1], do_rgb[-1] = True, True, True, True
    out = host(batch.enhance_batch(dev(imgs), do_blur, do_rgb, random))
    state = random.getstate()
    random.seed(17)
    L.sample_pairing_flags(n)
    for i in range(n):
        assert np.array_equal(out[i], L.enhance_image(imgs[i], bool(do_blur[i]), bool(do_rgb[i]))), i
    assert random.getstate() == state


def test_enhance_blur_generic_radius_and_odd_rows():
    """GaussianBlur outside the fast path of the vertical pass: box radius > 3 (radius 9) and rows whose byte length is not a
    multiple of four — against the integer model of libImaging's box blur (itself proven equal to Pillow)."""
    from oracle import intmodels as M
    r = np.random.default_rng(53)
    for (n, H, W), rad in (((2, 64, 48), 9.0), ((2, 40, 33), 2.2), ((1, 37, 50), 0.6)):
        imgs = r.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
        items = np.zeros(n, dtype=batch.ENH_ITEM_DTYPE)
        items["brightness"] = items["contrast"] = items["color"] = 1.0
        radius, ww, fw = M.box_blur_params(M.gaussian_box_radius(rad))
        items["blur"], items["box_radius"], items["ww"], items["fw"] = 1, radius, ww, fw
        luts = np.tile(np.arange(256, dtype=np.uint8), (n, 3, 1))
        out = host(batch.enhance_batch(dev(imgs), None, None, plan=(items, luts)))
        for i in range(n):
            assert np.array_equal(out[i], M.gaussian_blur(imgs[i], rad)), (H, W, rad)


def test_enhance_image_drop_in_equals_reference_golden(tmp_path):
    """The file-based drop-in against outputs of the REAL reference function (tests/golden/tranfo_*.npz)."""
    from conftest import load_golden
    from PIL import Image as PILImage
    from image_processor_pipeline_b200.transforms import tranfo
    for k in range(4):
        g = load_golden(f"tranfo_{k}")
        d_in, d_out = tmp_path / f"in{k}", tmp_path / f"out{k}"
        d_in.mkdir(); d_out.mkdir()
        PILImage.fromarray(g["input"], "RGB").save(d_in / "img.png")
        random.seed(g["meta"]["seed"])
        res = tranfo.enhance_image(d_in / "img.png", g["meta"]["apply_blur"], g["meta"]["apply_rgb"], [d_out])
        assert res.name == g["meta"]["name"]
        assert np.array_equal(np.asarray(PILImage.open(res).convert("RGB")), g["out"])
