"""T3 integer models == the installed cv2 / Pillow (T2), bit for bit.  CPU only."""
import numpy as np
import cv2
import pytest
from PIL import Image

from oracle import intmodels as M, libcalls as L

rng = np.random.default_rng(7)


@pytest.mark.parametrize("wh", [(64, 48), (333, 517), (640, 360)])
@pytest.mark.parametrize("angle", [17.3, 45.0, 90.0, 180.0, 270.0, 359.0, 1.0, 123.456, 200.2, 360.0, -30.0])
def test_rotate_nearest_model(wh, angle):
    w, h = wh
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    a, b = L.rotate_expand(img, angle), M.rotate_nearest(img, angle)
    assert a.shape == b.shape and np.array_equal(a, b)


def test_rotate_nearest_model_1080p():
    img = rng.integers(0, 256, (1080, 1920, 4), dtype=np.uint8)
    assert np.array_equal(L.rotate_expand(img, 17.3), M.rotate_nearest(img, 17.3))


@pytest.mark.parametrize("angle", [17.3, 45.0, 123.456, 200.2])
def test_rotate_bilinear_model(angle):
    img = rng.integers(0, 256, (217, 333, 4), dtype=np.uint8)
    a = L.rotate_expand(img, angle, Image.Resampling.BILINEAR)
    assert np.array_equal(a, M.rotate_bilinear(img, angle))


def test_alpha_bbox_model():
    img = np.zeros((40, 50, 4), np.uint8)
    assert M.alpha_bbox(img) is None and Image.fromarray(img, "RGBA").getbbox() is None
    img[5:9, 7:30, 3] = 3
    img[20, 45, 3] = 1
    img[0, 0, 0] = 200        # colour without alpha must not count
    assert M.alpha_bbox(img) == Image.fromarray(img, "RGBA").getbbox()
    rgb = np.zeros((10, 12, 3), np.uint8); rgb[3, 4, 1] = 1
    assert M.alpha_bbox(rgb) == Image.fromarray(rgb, "RGB").getbbox()


@pytest.mark.parametrize("pos", [(0, 0), (10, 20), (70, 60), (-5, -7), (100, 90), (119, 99), (200, 5)])
def test_paste_model(pos):
    bg = rng.integers(0, 256, (100, 120, 3), dtype=np.uint8)
    sp = rng.integers(0, 256, (40, 50, 4), dtype=np.uint8)
    assert np.array_equal(L.paste(bg, sp, pos), M.paste_rgba_on_rgb(bg, sp, pos))


@pytest.mark.parametrize("dims", [(80, 60, 48, 36), (50, 40, 75, 60), (200, 150, 31, 23), (64, 64, 64, 32),
                                  (33, 77, 33, 20), (512, 384, 300, 225), (7, 5, 1, 1), (3, 3, 20, 30)])
def test_resize_lanczos_model(dims):
    w, h, nw, nh = dims
    sp = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    assert np.array_equal(L.resize_lanczos(sp, (nw, nh)), M.resize_lanczos_rgba(sp, (nw, nh)))


def test_premultiply_roundtrip_model():
    c = np.arange(256, dtype=np.uint8)
    img = np.zeros((256, 256, 4), np.uint8)
    img[..., 0] = c[None, :]; img[..., 1] = 255 - c[None, :]; img[..., 2] = 77; img[..., 3] = c[:, None]
    pm = np.asarray(Image.fromarray(img, "RGBA").convert("RGBa"))
    assert np.array_equal(pm, M.premultiply(img))
    back = np.asarray(Image.frombuffer("RGBa", (256, 256), pm.tobytes(), "raw", "RGBa", 0, 1).convert("RGBA"))
    assert np.array_equal(back, M.unpremultiply(pm))


def test_bgr2hsv_model_all_colours():
    c = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([c & 255, (c >> 8) & 255, (c >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV), M.bgr2hsv(bgr))


def test_inrange_bound_rounding():
    vals = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    for lo, hi in [(76.5, 152.5), (77.5, 229.5), (0.5, 1.5), (2.5, 3.5), (-3.2, 300.7), (254.5, 255.5)]:
        a = cv2.inRange(vals, np.array([lo, lo, lo]), np.array([hi, hi, hi]))
        l3, h3 = M.inrange_bounds_u8([lo] * 3, [hi] * 3)
        b = ((vals[..., 0].astype(int) >= l3[0]) & (vals[..., 0].astype(int) <= h3[0])).astype(np.uint8) * 255
        assert np.array_equal(a, b), (lo, hi)


FILTERS = [(0, 0, 0, 180, 255, 150), (15, 60, 200, 35, 255, 255),
           (15, 30 * 2.55, 55 * 2.55, 30, 60 * 2.55, 80 * 2.55), (15, 60 * 2.55, 60 * 2.55, 30, 75 * 2.55, 90 * 2.55)]


def test_color_masks_model():
    img = rng.integers(0, 256, (64, 80, 3), dtype=np.uint8)
    zones = [(0, 0, 0, 0), (3, 4, 5, 6), None, (10, 0, 0, 20)]
    assert np.array_equal(L.color_masks_to_alpha(img, FILTERS, zones), M.color_masks_to_alpha(img, FILTERS, zones))
    assert np.array_equal(L.color_masks_to_alpha(img, FILTERS), M.color_masks_to_alpha(img, FILTERS))


def test_keep_largest_model_random_masks():
    r = np.random.default_rng(3)
    for trial in range(80):
        H, W = r.integers(3, 70, 2)
        p = r.choice([0.2, 0.4, 0.5, 0.6])
        alpha = (r.random((H, W)) < p).astype(np.uint8) * r.integers(2, 256, (H, W)).astype(np.uint8)
        alpha[r.random((H, W)) < 0.05] = 1
        if alpha.max() <= 1:
            continue
        img = np.dstack([r.integers(0, 256, (H, W, 3), dtype=np.uint8), alpha])
        _, cleaned, bbox, area = L.keep_largest_component(img, with_info=True)
        out, bb, ar = M.keep_largest_alpha(alpha)
        assert np.array_equal(cleaned[..., 3], out), trial
        assert tuple(bbox) == bb and ar == area


def test_keep_largest_model_tie_break_block_raster():
    # two 2-pixel components; pixel-raster order and 2x2-block-raster order disagree
    alpha = np.zeros((4, 8), np.uint8)
    alpha[1, 6] = alpha[1, 7] = 200     # block key (0, 3)
    alpha[2, 0] = alpha[3, 0] = 200     # block key (1, 0)
    alpha[0, 2] = 50                    # single pixel, block key (0, 1)
    img = np.dstack([np.zeros((4, 8, 3), np.uint8), alpha])
    _, cleaned, bbox, area = L.keep_largest_component(img, with_info=True)
    out, bb, ar = M.keep_largest_alpha(alpha)
    assert np.array_equal(cleaned[..., 3], out) and tuple(bbox) == bb and area == ar == 2


# ------------------------------------------------------------------ tranfo.py (photometric jitter)
@pytest.mark.parametrize("shape", [(37, 53, 3), (5, 200, 3), (64, 7, 3), (120, 90, 3)])
def test_enhance_models_equal_pillow(shape):
    from PIL import ImageEnhance, ImageFilter
    img = np.random.default_rng(3).integers(0, 256, shape, dtype=np.uint8)
    im = Image.fromarray(img, "RGB")
    for f in (0.7, 0.83, 1.0, 1.17, 1.3, 0.0, 2.5):
        assert np.array_equal(np.asarray(ImageEnhance.Brightness(im).enhance(f)), M.enhance_brightness(img, f))
        assert np.array_equal(np.asarray(ImageEnhance.Contrast(im).enhance(f)), M.enhance_contrast(img, f))
        assert np.array_equal(np.asarray(ImageEnhance.Color(im).enhance(f)), M.enhance_color(img, f))
    for rad in (0.5, 0.51, 0.77, 1.0, 1.5, 2.0, 2.37, 2.9, 3.0, 0.05, 7.3):
        assert np.array_equal(np.asarray(im.filter(ImageFilter.GaussianBlur(rad))), M.gaussian_blur(img, rad)), rad
    fac = np.random.default_rng(4).uniform(0.75, 1.25, 256)
    lut = M.point_lut(fac)
    ref = np.asarray(im.split()[0].point(lambda p: max(0, min(255, p * fac[p]))))
    assert np.array_equal(ref, lut[img[..., 0]])


def test_enhance_oracle_draw_order_matches_models():
    """oracle.libcalls.enhance_image (the Pillow calls with draws from `random`) == the integer models fed with the same draws,
    in the same order: 3 uniforms, optional blur radius, optional 3 x 256 LUT factors."""
    import random
    img = np.random.default_rng(5).integers(0, 256, (48, 70, 3), dtype=np.uint8)
    for blur, rgb in ((False, False), (True, False), (False, True), (True, True)):
        random.seed(9)
        ref = L.enhance_image(img, blur, rgb)
        state = random.getstate()
        random.seed(9)
        out = M.enhance_brightness(img, random.uniform(0.7, 1.3))
        out = M.enhance_contrast(out, random.uniform(0.7, 1.3))
        out = M.enhance_color(out, random.uniform(0.7, 1.3))
        if blur:
            out = M.gaussian_blur(out, random.uniform(0.5, 3))
        if rgb:
            luts = [M.point_lut([random.uniform(0.75, 1.25) for _ in range(256)]) for _ in range(3)]
            out = np.stack([luts[c][out[..., c]] for c in range(3)], axis=2)
        assert random.getstate() == state
        assert np.array_equal(ref, out)
