"""T3 oracle: pure-numpy integer restatements of the third-party arithmetic.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  These document, operation by
operation, what each CUDA kernel must compute.  The arithmetic lives in
un-vendored dependencies of the reference (no version pins anywhere in
/root/reference): opencv-python (container: 4.13.0) and Pillow (container:
12.2.0).  Each model below restates the published algorithm of that library
and is pinned bit-exact against the installed library by
``tests/test_oracle_models.py``; the reference call site is cited per function.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

# ----------------------------------------------------------------------------------
# Pillow Image.rotate(expand=True) — reference transforms/rotations.py:96
# ----------------------------------------------------------------------------------


def rotate_matrix(w: int, h: int, angle: float):
    """Pillow Image.rotate matrix + expanded canvas (PIL/Image.py:2509-2585).

    Returns (kind, matrix6, (nw, nh)) with kind in {"copy", "rot180", "rot90",
    "rot270", "affine"} mirroring the fast paths of PIL/Image.py:2509-2521."""
    angle = angle % 360.0
    if angle == 0:
        return "copy", None, (w, h)
    if angle == 180:
        return "rot180", None, (w, h)
    if angle in (90, 270):
        return ("rot90" if angle == 90 else "rot270"), None, (h, w)
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0,
         round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]

    def tf(x, y, m):
        return m[0] * x + m[1] * y + m[2], m[3] * x + m[4] * y + m[5]

    m[2], m[5] = tf(-cx, -cy, m)
    m[2] += cx
    m[5] += cy
    xx, yy = [], []
    for x, y in ((0, 0), (w, 0), (w, h), (0, h)):
        tx, ty = tf(x, y, m)
        xx.append(tx)
        yy.append(ty)
    nw = math.ceil(max(xx)) - math.floor(min(xx))
    nh = math.ceil(max(yy)) - math.floor(min(yy))
    m[2], m[5] = tf(-(nw - w) / 2.0, -(nh - h) / 2.0, m)
    return "affine", m, (nw, nh)


def _fix(v: float) -> int:
    """libImaging Geometry.c FIX(): FLOOR(v*65536 + 0.5), FLOOR = trunc for v>=0 else floor."""
    t = v * 65536.0 + 0.5
    return int(t) if t >= 0.0 else int(math.floor(t))


def affine_fixed_coeffs(m) -> Tuple[int, int, int, int, int, int]:
    """libImaging Geometry.c affine_fixed(): 16.16 fixed-point coefficients."""
    a0 = _fix(m[0])
    a1 = _fix(m[1])
    a3 = _fix(m[3])
    a4 = _fix(m[4])
    a2 = _fix(m[2] + m[0] * 0.5 + m[1] * 0.5)
    a5 = _fix(m[5] + m[3] * 0.5 + m[4] * 0.5)
    return a0, a1, a2, a3, a4, a5


def affine_fixed_ok(m, nw: int, nh: int) -> bool:
    """libImaging Geometry.c check_fixed() on the four output corners."""
    def ok(x, y):
        return (abs(x * m[0] + y * m[1] + m[2]) < 32768.0 and
                abs(x * m[3] + y * m[4] + m[5]) < 32768.0)
    return ok(0, 0) and ok(nw, nh) and ok(0, nh) and ok(nw, 0)


def rotate_nearest(rgba: np.ndarray, angle: float) -> np.ndarray:
    """Integer model of ``Image.rotate(angle, expand=True)`` (NEAREST) on RGBA."""
    h, w = rgba.shape[:2]
    kind, m, (nw, nh) = rotate_matrix(w, h, angle)
    if kind == "copy":
        return rgba.copy()
    if kind == "rot180":
        return rgba[::-1, ::-1].copy()
    if kind == "rot90":      # Transpose.ROTATE_90: counter-clockwise
        return np.ascontiguousarray(np.rot90(rgba, 1))
    if kind == "rot270":
        return np.ascontiguousarray(np.rot90(rgba, 3))
    assert affine_fixed_ok(m, nw, nh)
    a0, a1, a2, a3, a4, a5 = affine_fixed_coeffs(m)
    ys = np.arange(nh, dtype=np.int64)[:, None]
    xs = np.arange(nw, dtype=np.int64)[None, :]
    xin = (a2 + ys * a1 + xs * a0) >> 16
    yin = (a5 + ys * a4 + xs * a3) >> 16
    valid = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros((nh, nw, rgba.shape[2]), dtype=np.uint8)
    out[valid] = rgba[yin[valid], xin[valid]]
    return out


def alpha_bbox(img: np.ndarray):
    """Pillow getbbox() (alpha_only=True for RGBA/LA, any-channel otherwise;
    PIL/Image.py:1480-1497).  Returns (x0, y0, x1, y1) exclusive or None."""
    if img.ndim == 3 and img.shape[2] in (2, 4):
        nz = img[:, :, -1] != 0
    elif img.ndim == 3:
        nz = img.any(axis=2)
    else:
        nz = img != 0
    rows = np.flatnonzero(nz.any(axis=1))
    if rows.size == 0:
        return None
    cols = np.flatnonzero(nz.any(axis=0))
    return int(cols[0]), int(rows[0]), int(cols[-1]) + 1, int(rows[-1]) + 1


# ---- bilinear rotate (extra mode; the reference itself uses NEAREST) -----------------


def premultiply(rgba: np.ndarray) -> np.ndarray:
    """libImaging Convert.c rgbA2rgba(): MULDIV255 of the colour bands by alpha."""
    a = rgba[..., 3:4].astype(np.uint32)
    t = rgba[..., :3].astype(np.uint32) * a + 128
    out = rgba.copy()
    out[..., :3] = (((t >> 8) + t) >> 8).astype(np.uint8)
    return out


def unpremultiply(rgba_pm: np.ndarray) -> np.ndarray:
    """libImaging Convert.c rgba2rgbA(): CLIP8(255*c/alpha), copy when alpha in {0,255}."""
    a = rgba_pm[..., 3:4].astype(np.uint32)
    c = rgba_pm[..., :3].astype(np.uint32)
    safe = np.where(a == 0, 1, a)
    q = np.minimum(255, (255 * c) // safe)
    res = np.where((a == 0) | (a == 255), c, q).astype(np.uint8)
    out = rgba_pm.copy()
    out[..., :3] = res
    return out


def rotate_bilinear(rgba: np.ndarray, angle: float) -> np.ndarray:
    """Model of ``Image.rotate(angle, resample=BILINEAR, expand=True)`` on RGBA:
    RGBA->RGBa, libImaging ImagingGenericTransform + bilinear_filter32RGB
    (double precision, clamp-to-edge, truncation to u8), RGBa->RGBA."""
    h, w = rgba.shape[:2]
    kind, m, (nw, nh) = rotate_matrix(w, h, angle)
    if kind != "affine":
        return rotate_nearest(rgba, angle)
    src = premultiply(rgba).astype(np.float64)
    ys = np.arange(nh, dtype=np.float64)[:, None] + 0.5
    xs = np.arange(nw, dtype=np.float64)[None, :] + 0.5
    xin = m[0] * xs + m[1] * ys + m[2]
    yin = m[3] * xs + m[4] * ys + m[5]
    valid = (xin >= 0.0) & (xin < w) & (yin >= 0.0) & (yin < h)
    xin = xin - 0.5
    yin = yin - 0.5
    x = np.floor(xin).astype(np.int64)
    y = np.floor(yin).astype(np.int64)
    dx = xin - x
    dy = yin - y
    x0 = np.clip(x, 0, w - 1)
    x1 = np.clip(x + 1, 0, w - 1)
    y0 = np.clip(y, 0, h - 1)
    y1 = np.clip(y + 1, 0, h - 1)
    out = np.zeros((nh, nw, 4), dtype=np.uint8)
    for c in range(4):
        ch = src[..., c]
        v1 = ch[y0, x0] + (ch[y0, x1] - ch[y0, x0]) * dx
        v2 = ch[y1, x0] + (ch[y1, x1] - ch[y1, x0]) * dx
        v = v1 + (v2 - v1) * dy
        out[..., c] = np.where(valid, v.astype(np.int64).astype(np.uint8), 0)
    return unpremultiply(out)


# ----------------------------------------------------------------------------------
# Pillow paste with RGBA mask — reference transforms/overlays.py:138-139
# ----------------------------------------------------------------------------------


def div255(t: np.ndarray) -> np.ndarray:
    """libImaging ImagingUtils.h DIV255: t+=128; ((t>>8)+t)>>8."""
    t = t + 128
    return ((t >> 8) + t) >> 8


def paste_rgba_on_rgb(bg_rgb: np.ndarray, sprite_rgba: np.ndarray, pos: Tuple[int, int]) -> np.ndarray:
    """libImaging Paste.c paste_mask_RGBA: out = DIV255(bg*(255-a) + spr*a), clipped to bg."""
    out = bg_rgb.copy()
    H, W = bg_rgb.shape[:2]
    sh, sw = sprite_rgba.shape[:2]
    px, py = pos
    x0, y0 = max(px, 0), max(py, 0)
    x1, y1 = min(px + sw, W), min(py + sh, H)
    if x1 <= x0 or y1 <= y0:
        return out
    spr = sprite_rgba[y0 - py:y1 - py, x0 - px:x1 - px].astype(np.uint32)
    a = spr[..., 3:4]
    dst = out[y0:y1, x0:x1].astype(np.uint32)
    out[y0:y1, x0:x1] = div255(dst * (255 - a) + spr[..., :3] * a).astype(np.uint8)
    return out


# ----------------------------------------------------------------------------------
# Pillow resize(LANCZOS) — reference transforms/overlays.py:129
# ----------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def lanczos_coeffs(in_size: int, out_size: int):
    """libImaging Resample.c precompute_coeffs + normalize_coeffs_8bpc for the
    full-image box (0, in_size).  Returns (ksize, bounds[out,2] int32, kk[out,ksize] int32)."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ws = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for wv in ws:
            ww += wv
        for x in range(xmax):
            wv = ws[x] / ww if ww != 0.0 else ws[x]
            if wv < 0:
                kk[xx, x] = int(-0.5 + wv * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + wv * (1 << PRECISION_BITS))
        bounds[xx, 0] = xmin
        bounds[xx, 1] = xmax
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_pass_h(src: np.ndarray, out_w: int) -> np.ndarray:
    """libImaging ImagingResampleHorizontal_8bpc on a 4-band u8 image."""
    h, w = src.shape[:2]
    _, bounds, kk = lanczos_coeffs(w, out_w)
    out = np.empty((h, out_w, src.shape[2]), dtype=np.uint8)
    s = src.astype(np.int64)
    for xx in range(out_w):
        xmin, n = bounds[xx]
        acc = (s[:, xmin:xmin + n, :] * kk[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1)
        out[:, xx, :] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
    return out


def resample_pass_v(src: np.ndarray, out_h: int) -> np.ndarray:
    """libImaging ImagingResampleVertical_8bpc on a 4-band u8 image."""
    h, w = src.shape[:2]
    _, bounds, kk = lanczos_coeffs(h, out_h)
    out = np.empty((out_h, w, src.shape[2]), dtype=np.uint8)
    s = src.astype(np.int64)
    for yy in range(out_h):
        ymin, n = bounds[yy]
        acc = (s[ymin:ymin + n, :, :] * kk[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0)
        out[yy] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
    return out


def resize_lanczos_rgba(rgba: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """Model of ``Image.resize(size, LANCZOS)`` on RGBA (PIL/Image.py resize:
    same-size -> copy; RGBA -> RGBa -> two integer passes -> RGBA)."""
    nw, nh = size
    h, w = rgba.shape[:2]
    if (w, h) == (nw, nh):
        return rgba.copy()
    cur = premultiply(rgba)
    if nw != w:
        cur = resample_pass_h(cur, nw)
    if nh != h:
        cur = resample_pass_v(cur, nh)
    return unpremultiply(cur)


# ----------------------------------------------------------------------------------
# OpenCV BGR2HSV (8-bit, H in [0,180)) + inRange — reference filtres_liste.py:90-117
# ----------------------------------------------------------------------------------
HSV_SHIFT = 12


def _cv_round(x: float) -> int:
    """cvRound: round half to even."""
    return int(np.rint(x))


def hsv_tables():
    sdiv = np.zeros(256, dtype=np.int32)
    hdiv180 = np.zeros(256, dtype=np.int32)
    for i in range(1, 256):
        sdiv[i] = _cv_round((255 << HSV_SHIFT) / (1.0 * i))
        hdiv180[i] = _cv_round((180 << HSV_SHIFT) / (6.0 * i))
    return sdiv, hdiv180


_SDIV, _HDIV180 = hsv_tables()


def bgr2hsv(bgr: np.ndarray) -> np.ndarray:
    """OpenCV imgproc color_hsv RGB2HSV_b (hrange=180) integer model."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    s = (diff * _SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV180[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


def inrange_bounds_u8(lo, hi):
    """cv2.inRange scalar bounds on a CV_8U image: each bound is rounded half-to-even
    to int; the test is lo<=px<=hi on integers (so an out-of-range bound simply
    never/always matches).  Returns int arrays (lo3, hi3)."""
    lo_i = [int(np.rint(float(x))) for x in lo]
    hi_i = [int(np.rint(float(x))) for x in hi]
    return np.array(lo_i, dtype=np.int64), np.array(hi_i, dtype=np.int64)


def color_masks_to_alpha(bgr: np.ndarray, ranges, zones=None) -> np.ndarray:
    """Integer model of filtres_liste.py:90-134: alpha = ~OR_k(inRange_k & zone_k)."""
    H, W = bgr.shape[:2]
    hsv = bgr2hsv(bgr).astype(np.int64)
    excl = np.zeros((H, W), dtype=bool)
    zones = zones or [None] * len(ranges)
    for filt, zone in zip(ranges, zones):
        lo, hi = inrange_bounds_u8(filt[:3], filt[3:])
        m = np.all((hsv >= lo) & (hsv <= hi), axis=-1)
        t, btm, l, r = zone if zone else (0, 0, 0, 0)
        zm = np.zeros((H, W), dtype=bool)
        zm[t:H - btm, l:W - r] = True
        excl |= m & zm
    out = np.empty((H, W, 4), dtype=np.uint8)
    out[..., :3] = bgr
    out[..., 3] = np.where(excl, 0, 255)
    return out


# ----------------------------------------------------------------------------------
# 8-connected components, keep largest — reference pixels_isolés.py:32-55
# ----------------------------------------------------------------------------------


def label_components_8(binary: np.ndarray):
    """Plain union-find 8-connected labelling.  Returns (labels int32 with 0 =
    background and components numbered 1.. in ascending order of their minimum
    2x2-block raster key ((y//2)*ceil(W/2) + x//2), areas[int64], keys)."""
    H, W = binary.shape
    fg = binary != 0
    idx = np.arange(H * W, dtype=np.int64).reshape(H, W)
    parent = np.arange(H * W, dtype=np.int64)

    def find(a):
        root = a
        while parent[root] != root:
            root = parent[root]
        while parent[a] != root:
            parent[a], a = root, parent[a]
        return root

    ys, xs = np.nonzero(fg)
    for y, x in zip(ys.tolist(), xs.tolist()):
        for dy, dx in ((0, -1), (-1, -1), (-1, 0), (-1, 1)):
            yy, xx = y + dy, x + dx
            if 0 <= yy < H and 0 <= xx < W and fg[yy, xx]:
                ra, rb = find(idx[y, x]), find(idx[yy, xx])
                if ra != rb:
                    if ra < rb:
                        parent[rb] = ra
                    else:
                        parent[ra] = rb
    roots = np.array([find(i) for i in idx[fg].tolist()], dtype=np.int64)
    bw = (W + 1) // 2
    keys_px = (ys // 2) * bw + (xs // 2)
    uniq, inv = np.unique(roots, return_inverse=True)
    comp_key = np.full(uniq.size, np.iinfo(np.int64).max)
    np.minimum.at(comp_key, inv, keys_px)
    order = np.argsort(comp_key, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    labels = np.zeros((H, W), dtype=np.int32)
    labels[fg] = rank[inv] + 1
    areas = np.bincount(rank[inv], minlength=uniq.size).astype(np.int64)
    return labels, areas, comp_key[order]


def keep_largest_alpha(alpha: np.ndarray):
    """Integer model of pixels_isolés.py:32-55 on the alpha plane.

    Returns (alpha_out, bbox(x, y, w, h) of alpha_out != 0 or (0,0,0,0), area)."""
    binary = alpha > 1
    labels, areas, _ = label_components_8(binary)
    out = alpha.copy()
    if areas.size == 0:
        best_area = 0
    else:
        best = int(np.argmax(areas))      # first maximum = smallest block-raster key
        best_area = int(areas[best])
        out[labels != best + 1] = 0
    nz = out != 0
    if not nz.any():
        return out, (0, 0, 0, 0), best_area
    rows = np.flatnonzero(nz.any(axis=1))
    cols = np.flatnonzero(nz.any(axis=0))
    return out, (int(cols[0]), int(rows[0]), int(cols[-1] - cols[0] + 1), int(rows[-1] - rows[0] + 1)), best_area


# --------------------------------------------------------------------------- photometric jitter (tranfo.py -> Pillow)
def luma_u8(rgb: np.ndarray) -> np.ndarray:
    """Pillow convert("L") (libImaging Convert.c rgb2l): (R*19595 + G*38470 + B*7471 + 0x8000) >> 16."""
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend_u8(a: np.ndarray, b: np.ndarray, alpha: float) -> np.ndarray:
    """libImaging ImagingBlend (Blend.c): out = (UINT8)(a + alpha * (b - a)) in SINGLE precision, truncation;
    outside [0, 1] the result is clipped to [0, 255] first.  ImageEnhance.*.enhance(f) = blend(degenerate, image, f)."""
    alpha = np.float32(alpha)
    a_i, b_i = a.astype(np.int32), b.astype(np.int32)
    t = (a_i.astype(np.float32) + alpha * (b_i - a_i).astype(np.float32)).astype(np.float32)
    if 0 <= alpha <= 1:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def enhance_brightness(rgb, f):
    return blend_u8(np.zeros_like(rgb), rgb, f)                       # degenerate = black (ImageEnhance.Brightness)


def contrast_mean(rgb) -> int:
    lum = luma_u8(rgb)
    return int(int(lum.astype(np.int64).sum()) / lum.size + 0.5)      # int(ImageStat.Stat(L).mean[0] + 0.5)


def enhance_contrast(rgb, f):
    return blend_u8(np.full_like(rgb, contrast_mean(rgb)), rgb, f)    # degenerate = mean gray (ImageEnhance.Contrast)


def enhance_color(rgb, f):
    return blend_u8(np.repeat(luma_u8(rgb)[..., None], 3, axis=2), rgb, f)    # degenerate = gray image (ImageEnhance.Color)


def gaussian_box_radius(radius: float, passes: int = 3) -> np.float32:
    """libImaging BoxBlur.c _gaussian_blur_radius: the box radius whose `passes` box blurs approximate a gaussian of
    that sigma — single-precision variables, double-precision sqrt / floor as the C expression promotes them."""
    import math
    f32 = np.float32
    radius = f32(radius)
    sigma2 = f32(radius * radius / f32(passes))
    L = f32(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = f32(math.floor((float(L) - 1.0) / 2.0))
    a = f32(f32(2 * l + 1) * f32(f32(l * f32(l + 1)) - f32(3 * sigma2)))
    a = f32(a / f32(6 * f32(sigma2 - f32(f32(l + 1) * f32(l + 1)))))
    return f32(l + a)


def box_blur_params(float_radius):
    """ImagingHorizontalBoxBlur: integer radius, 8.24 fixed-point weight of a full pixel (ww) and of the two far pixels (fw)."""
    f32 = np.float32
    radius = int(float_radius)
    ww = int(f32(16777216) / f32(f32(f32(float_radius) * f32(2)) + f32(1)))
    fw = ((1 << 24) - (radius * 2 + 1) * ww) // 2
    return radius, ww, fw


def box_blur_rows(a: np.ndarray, float_radius) -> np.ndarray:
    """One horizontal box-blur pass (ImagingLineBoxBlur32/8): window sum over [x-r, x+r] with clamped edges times ww plus the
    two far pixels times fw, (v + 2^23) >> 24 in unsigned 32-bit arithmetic."""
    H, W = a.shape[:2]
    radius, ww, fw = box_blur_params(float_radius)
    idx = np.arange(W)
    ai = a.astype(np.int64)
    acc = np.zeros_like(ai)
    for d in range(-radius, radius + 1):
        acc += ai[:, np.clip(idx + d, 0, W - 1)]
    far = ai[:, np.clip(idx - radius - 1, 0, W - 1)] + ai[:, np.clip(idx + radius + 1, 0, W - 1)]
    bulk = (acc * ww + far * fw) & 0xffffffff
    return (((bulk + (1 << 23)) & 0xffffffff) >> 24).astype(np.uint8)


def gaussian_blur(rgb: np.ndarray, radius: float) -> np.ndarray:
    """ImageFilter.GaussianBlur(radius) = ImagingGaussianBlur(passes=3): three horizontal box passes (each rounded to 8
    bits), transpose, three more, transpose back."""
    br = gaussian_box_radius(radius)
    out = rgb
    if br != 0:
        for _ in range(3):
            out = box_blur_rows(out, br)
        t = np.ascontiguousarray(out.transpose(1, 0, 2))
        for _ in range(3):
            t = box_blur_rows(t, br)
        out = np.ascontiguousarray(t.transpose(1, 0, 2))
    return out


def point_lut(factors) -> np.ndarray:
    """Image.point(lambda p: max(0, min(255, p * f_p))) with one factor per table entry: round-half-even of the clamped
    product (Image.point rounds float tables with Python's round)."""
    return np.array([round(max(0, min(255, p * f))) for p, f in enumerate(factors)], dtype=np.uint8)
