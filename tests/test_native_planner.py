"""The C++ host planner (planner.cu) against its readable Python twin (chain.py) and against
Python's own `random`: same draws, same records, same labels, same generator state afterwards.
CPU only (the planner is host code)."""
import ctypes as C
import math
import random

import numpy as np
import pytest

from image_processor_pipeline_b200 import _native, chain
from test_host_logic import _fake_stage_a


def _fake_b1_arr(rots):
    out = np.zeros((len(rots), 4), np.int32)
    r = np.random.default_rng(len(rots))
    for i, rec in enumerate(rots):
        cw, ch = int(rec["canvas_w"]), int(rec["canvas_h"])
        if i % 7 == 3:
            out[i] = (2 ** 31 - 1, 2 ** 31 - 1, 0, 0)               # empty bbox -> un-cropped canvas
        else:
            x0, y0 = int(r.integers(0, max(1, cw // 8))), int(r.integers(0, max(1, ch // 8)))
            out[i] = (x0, y0, cw - int(r.integers(0, max(1, cw // 8))), ch - int(r.integers(0, max(1, ch // 8))))
    return out


CASES = [
    dict(num_rotations=2, include_original=True, pool=["o", "h", "v", "hv"], choose_random=2),
    dict(num_rotations=1, include_original=False, pool=["o", "h", "v", "hv"], choose_random=1, sym_include_original=False),
    dict(num_rotations=3, include_original=False, angle_min=80.0, angle_max=100.0, pool=["hv"], scale_min=0.4, scale_max=0.9, yolo_class_id=7),
    dict(num_rotations=1, include_original=True, pool=["h", "v"], choose_random=0, sym_include_original=True, scale_min=0.9, scale_max=1.6),
    dict(num_rotations=4, include_original=False, pool=None, choose_random=None, scale_min=0.01, scale_max=0.05),
]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("bg", [(1920, 1080, 3), (220, 300, 2), (640, 640, 5)])
def test_native_planner_equals_python_planner(kw, bg):
    BW, BH, m = bg
    p = chain.ChainParams(color_ranges_to_exclude_hsv=[(0, 0, 0, 180, 255, 150)], **kw)
    info = _fake_stage_a(9, BW)
    info[:, 2] = np.minimum(info[:, 2], 700); info[:, 3] = np.minimum(info[:, 3], 600)
    # python planner
    rng_py = random.Random(42)
    rots = chain.plan_rotations(info, p, rng_py)
    rot_arr, _ = chain._rot_records(rots)
    bbox = _fake_b1_arr(rot_arr)
    finals = chain.plan_items(rots, bbox, p, rng_py, m, BW, BH)
    items_py, aux_py = chain._python_plan_to_arrays(rots, finals)
    # native planner on the same generator state
    rng_nat = random.Random(42)
    bridge = chain._RngBridge(rng_nat)
    pp = chain._plan_params(p, BW, BH, m)
    rots_n, angles_n, rot_index_n = chain.native_plan_rotations(info, pp, bridge)
    for name in ("frame", "cc_x", "cc_y", "cc_w", "cc_h", "kind", "a0", "a1", "a2", "a3", "a4", "a5", "canvas_w", "canvas_h", "tile_start"):
        assert np.array_equal(rots_n[name], rot_arr[name]), name
    assert [None if a != a else a for a in angles_n] == [r["angle"] for r in rots]
    assert list(rot_index_n) == [r["rot_index"] for r in rots]
    items_n, aux_n, labels_n = chain.native_plan_items(rots_n, rot_index_n, bbox, pp, bridge)
    bridge.commit()
    assert len(items_n) == len(items_py)
    for name in ("frame", "bg", "cc_x", "cc_y", "cc_w", "cc_h", "kind", "a0", "a1", "a2", "a3", "a4", "a5", "fit_x0", "fit_y0",
                 "spr_w", "spr_h", "flip_h", "flip_v", "new_w", "new_h", "pos_x", "pos_y", "sq_x0", "sq_y0", "sq_size"):
        assert np.array_equal(items_n[name], items_py[name]), name
    assert np.array_equal(aux_n["rot"], aux_py["rot"]) and np.array_equal(aux_n["sym"], aux_py["sym"])
    assert np.array_equal(aux_n["ratio"], aux_py["ratio"])
    assert labels_n == [f["label"] for f in finals]
    assert rng_nat.getstate() == rng_py.getstate()
    # layout produced natively == layout recomputed in numpy
    nat_layout = {k: items_n[k].copy() for k in ("h_ksize", "v_ksize", "hb_off", "hk_off", "vb_off", "vk_off", "gather_tile_start",
                                                 "h_tile_start", "v_tile_start", "tab_row_start", "rot_off", "tmp_off", "spr_off")}
    chain._layout(items_n)
    for k, v in nat_layout.items():
        assert np.array_equal(items_n[k], v), k


def test_native_rng_primitives_equal_python_random():
    """uniform / randint / sample draw-for-draw, through the planner's public behaviour: a pool of 4,
    choose 3 -> the symmetry order; randint bounds covering every bit length."""
    lib = _native.load()
    for seed in range(5):
        for w in (1, 2, 3, 5, 64, 1000, 4097):
            info = np.zeros((40, 8), np.int32); info[:, 2] = w; info[:, 3] = w
            p = chain.ChainParams(color_ranges_to_exclude_hsv=[(0, 0, 0, 1, 1, 1)], num_rotations=1, include_original=False,
                                  pool=["o", "h", "v", "hv"], choose_random=3, sym_include_original=False, scale_min=0.05, scale_max=0.9)
            a, b = random.Random(seed), random.Random(seed)
            rots = chain.plan_rotations(info, p, a)
            rot_arr, _ = chain._rot_records(rots)
            bbox = np.tile(np.array([[2 ** 31 - 1, 2 ** 31 - 1, 0, 0]], np.int32), (len(rots), 1))
            finals = chain.plan_items(rots, bbox, p, a, 3, 5000, 4000)
            bridge = chain._RngBridge(b)
            pp = chain._plan_params(p, 5000, 4000, 3)
            r_n, ang, ri = chain.native_plan_rotations(info, pp, bridge)
            it_n, aux_n, lab_n = chain.native_plan_items(r_n, ri, bbox, pp, bridge)
            bridge.commit()
            assert a.getstate() == b.getstate()
            assert lab_n == [f["label"] for f in finals]
            assert [chain.SYM_NAME[s] for s in aux_n["sym"]] == [f["sym"] for f in finals]


def test_hypot_and_round15_conventions():
    """std::hypot and printf/strtod rounding agree with math.hypot and round(x, 15) (spot check through the planner:
    canvas sizes and overlay sizes above depend on them); direct check of a few thousand values via Python only."""
    r = random.Random(1)
    for _ in range(2000):
        x = r.uniform(-1, 1)
        assert float(f"{x:.15f}") == round(x, 15)
    assert math.hypot(1920, 1080) == math.sqrt(1920 ** 2 + 1080 ** 2)


def test_exact_decimal_helpers_match_python():
    """round(x, 15) and "%.6f" + float() of the planner's integer fast paths vs Python, incl. exact ties,
    values next to ties, tiny / large magnitudes and float32-valued inputs (what labels hold)."""
    import ctypes as C
    import math
    import struct
    lib = _native.load()
    r = np.random.default_rng(3)
    xs = list(r.uniform(-1, 1, 20000)) + list(np.cos(r.uniform(0, 7, 5000))) + [0.0, 1.0, -1.0, 0.5, 1e-16, -3e-17, 2.5e-16, 0.1, 0.3,
                                                                                   1e-300, 5e-16, 15e-16 / 2, 0.9999999999999996]
    for k in range(2000):                                      # neighbours of 15-decimal ties
        t = (int(r.integers(0, 10 ** 15)) + 0.5) / 1e15
        xs += [t, math.nextafter(t, 0.0), math.nextafter(t, 2.0)]
    for x in xs:
        assert lib.ipp_host_round15(float(x)) == round(float(x), 15), x
    vs = list(r.uniform(0, 1, 20000).astype(np.float32).astype(np.float64)) + list(r.uniform(0, 1, 5000)) + [0.0, 1.0, 0.5, 1e-7, 5e-7, 4.9999999e-7,
                                                                                                            0.9999995, 0.99999949, 2.0000005, 1234.5678915, 1e-30]
    for k in range(2000):                                      # exact 6-decimal ties representable in binary: k / 64 + odd multiples of 2^-k
        vs += [(int(r.integers(0, 10 ** 6)) + 0.5) / 1e6, int(r.integers(0, 2 ** 20)) / 2 ** 20, int(r.integers(0, 2 ** 7)) / 2 ** 7]
    buf = C.create_string_buffer(64)
    parsed = C.c_double()
    for v in vs:
        n = lib.ipp_host_format6(float(v), buf, C.byref(parsed))
        want = "%.6f" % float(v)
        assert buf.value.decode() == want and n == len(want), (v, buf.value, want)
        assert parsed.value == float(want), v


def test_native_enhance_plan_equals_python_draws_and_pillow_models():
    """ipp_plan_enhance_host: the draws of tranfo.py:38-51 from the generator's own state (state equal afterwards), the
    GaussianBlur box parameters of libImaging and the Image.point tables — against Python's random + the integer models."""
    import random
    from image_processor_pipeline_b200 import batch
    from oracle import intmodels as M
    n = 23
    rng = random.Random(41)
    do_blur, do_rgb = batch.sample_flags(n, rng)
    assert do_blur.sum() == int(n * 0.3) and do_rgb.sum() == int(n * 0.3)
    ref = random.Random(41)
    files = list(range(n))
    rb = set(ref.sample(files, int(n * 0.3))); rr = set(ref.sample(files, int(n * 0.3)))
    assert [i in rb for i in files] == do_blur.tolist() and [i in rr for i in files] == do_rgb.tolist()
    items, luts, draws = batch.plan_enhance(n, do_blur, do_rgb, rng)
    for i in range(n):
        f = [ref.uniform(0.7, 1.3) for _ in range(3)]
        assert f == draws[i, :3].tolist()
        assert (np.float32(f[0]), np.float32(f[1]), np.float32(f[2])) == (items[i]["brightness"], items[i]["contrast"], items[i]["color"])
        if do_blur[i]:
            rad = ref.uniform(0.5, 3)
            assert rad == draws[i, 3]
            radius, ww, fw = M.box_blur_params(M.gaussian_box_radius(rad))
            assert (items[i]["blur"], items[i]["box_radius"], items[i]["ww"], items[i]["fw"]) == (1, radius, ww, fw)
        else:
            assert np.isnan(draws[i, 3]) and items[i]["blur"] == 0
        if do_rgb[i]:
            for c in range(3):
                assert np.array_equal(luts[i, c], M.point_lut([ref.uniform(0.75, 1.25) for _ in range(256)]))
        else:
            assert np.array_equal(luts[i], np.tile(np.arange(256, dtype=np.uint8), (3, 1)))
    assert rng.getstate() == ref.getstate()


def test_native_hypot_and_overlay_size_equal_python():
    """The planner's math.hypot (CPython's vector_norm, not libm's hypot) and the overlay size built on it (overlays.py:106-126)
    against Python, on random inputs and on the height-limited cases where the expression is mathematically an integer and the
    last bit of hypot decides int() (the fuzz case that exposed it: sprite 198 x 308 on a 319 x 105 background)."""
    import math
    import random
    from image_processor_pipeline_b200 import hostmath
    lib = _native.load()
    rnd = random.Random(5)
    for _ in range(20000):
        a, b = rnd.uniform(0, 5000), rnd.uniform(0, 5000)
        if rnd.random() < 0.3:
            a, b = float(int(a)), float(int(b))
        assert lib.ipp_host_hypot(a, b) == math.hypot(a, b)
    out = np.zeros(2, dtype=np.int32)
    cases = [(198, 308, 319, 105, 0.4329757405351776), (12, 564, 3595, 1968, 0.6696314906938445)]
    for _ in range(20000):
        cases.append((rnd.randint(1, 900), rnd.randint(1, 900), rnd.randint(8, 4000), rnd.randint(8, 2200), rnd.uniform(0.02, 1.2)))
    for ow, oh, bw, bh, ratio in cases:
        ok = lib.ipp_host_overlay_size(ow, oh, bw, bh, ratio, out.ctypes.data)
        nw, nh = hostmath.overlay_size(ow, oh, bw, bh, ratio)
        assert (int(out[0]), int(out[1])) == (nw, nh), (ow, oh, bw, bh, ratio)
        assert bool(ok) == (nw > 0 and nh > 0)
