"""Pin the oracle (T2 library-call restatements + T3 integer models) against outputs
of the REAL reference captured by tests/golden/make_golden.py.  CPU only."""
import random

import cv2
import numpy as np
import pytest

from conftest import load_golden
from oracle import intmodels as M, libcalls as L


@pytest.mark.parametrize("k", [0, 1, 2])
def test_symmetry_golden(k):
    g = load_golden(f"symmetry_{k}")
    random.seed(g["meta"]["seed"])
    opts = g["meta"]["opts"]
    syms = L.select_symmetries(opts.get("pool"), opts.get("choose_random"), opts.get("include_original", True))
    assert [f"img_{s}.png" for s in syms] == g["meta"]["names"]
    for i, out in enumerate(L.symmetries(g["input"], syms)):
        assert np.array_equal(out, g[f"out_{i}"])


@pytest.mark.parametrize("k", [0, 1])
def test_rotations_golden(k):
    g = load_golden(f"rotations_{k}")
    opts = g["meta"]["opts"]
    rgba = L.to_rgba(g["input"])
    i0 = 0
    if opts.get("include_original", True):
        assert np.array_equal(g["out_0"], rgba)
        i0 = 1
    for j, angle in enumerate(g["angles"]):
        t2, _ = L.rotate_and_fit(rgba, float(angle))
        assert np.array_equal(t2, g[f"out_{i0 + j}"])
        t3 = M.rotate_nearest(rgba, float(angle))
        x0, y0, x1, y1 = M.alpha_bbox(t3)
        assert np.array_equal(t3[y0:y1, x0:x1], g[f"out_{i0 + j}"])


@pytest.mark.parametrize("k", [0, 1, 2])
def test_crop_square_golden(k):
    g = load_golden(f"crop_square_{k}")
    data = np.loadtxt(g["meta"]["labels_in"].splitlines(), ndmin=2)
    cls, boxes = data[:, 0].astype(int), data[:, 1:5].astype(float)
    random.seed(g["meta"]["seed"])
    cropped, ncls, nboxes, _ = L.square_crop(g["input"], cls, boxes)
    assert np.array_equal(cropped, g["out_image"])
    assert L.format_label_rows(ncls, nboxes) == g["meta"]["labels_out"]


def test_crop_from_border_golden():
    g = load_golden("crop_from_border_0")
    dec = cv2.imdecode(g["input_jpg"], cv2.IMREAD_UNCHANGED)
    assert np.array_equal(dec, g["input_decoded"])
    crop = L.crop_from_border(dec, tuple(g["meta"]["margins"]))
    ok, enc = cv2.imencode(".jpg", crop)
    assert ok and np.array_equal(cv2.imdecode(enc, cv2.IMREAD_UNCHANGED), g["out_decoded"])


def test_fit_crop_golden():
    g = load_golden("fit_crop_0")
    assert np.array_equal(L.fit_crop(g["input"]), g["out"])
    x0, y0, x1, y1 = M.alpha_bbox(g["input"])
    assert np.array_equal(g["input"][y0:y1, x0:x1], g["out"])


@pytest.mark.parametrize("k", [0, 1, 2])
def test_overlays_golden(k):
    g = load_golden(f"overlays_{k}")
    random.seed(g["meta"]["seed"])
    comp, label, (nw, nh, px, py) = L.overlay_onto_background(g["sprite"], g["background"], **g["meta"]["opts"])
    assert np.array_equal(comp, g["out_image"])
    assert label == g["meta"]["label"]
    # T3: integer resample + DIV255 paste
    resized = M.resize_lanczos_rgba(g["sprite"], (nw, nh))
    assert np.array_equal(M.paste_rgba_on_rgb(g["background"], resized, (px, py)), g["out_image"])


@pytest.mark.parametrize("k", [0, 1, 2])
def test_pixels_isoles_golden(k):
    g = load_golden(f"pixels_isoles_{k}")
    assert np.array_equal(L.keep_largest_component(g["input"]), g["out"])
    alpha, (x, y, w, h), _ = M.keep_largest_alpha(g["input"][..., 3])
    model = g["input"].copy()
    model[..., 3] = alpha
    assert np.array_equal(model[y:y + h, x:x + w], g["out"])


@pytest.mark.parametrize("k", [0, 1, 2])
def test_filtres_liste_golden(k):
    g = load_golden(f"filtres_liste_{k}")
    m = g["meta"]
    filters = [tuple(f) for f in m["filters"]]
    zones = [tuple(z) if z else None for z in m["zones"]] if m["zones"] else None
    assert np.array_equal(L.color_masks_to_alpha(g["input"], filters, zones, m["gimp"]), g["out"])
    scaled = [L.rescale_filter(f, m["gimp"]) for f in filters]
    assert np.array_equal(M.color_masks_to_alpha(g["input"], scaled, zones), g["out"])


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_tranfo_golden(k):
    """transforms/tranfo.py:37-53 run for real (brightness / contrast / colour, optional blur, optional per-channel LUT)."""
    g = load_golden(f"tranfo_{k}")
    random.seed(g["meta"]["seed"])
    out = L.enhance_image(g["input"], g["meta"]["apply_blur"], g["meta"]["apply_rgb"])
    assert np.array_equal(out, g["out"])
