"""Host-side scalar math of the transform chain.

The reference computes these few numbers per image in Python (directly or inside
Pillow's Python layer) before any pixel is touched; the kernels take them as
explicit parameters.  Each function cites what it restates.  Everything here is
float64 / Python-int arithmetic in the reference's operation order, so the
results are the reference's results.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native


# --------------------------------------------------------------------------- rotation
@dataclass
class RotationPlan:
    kind: str                       # "affine" | "copy" | "rot180" | "rot90" | "rot270"
    size: Tuple[int, int]           # canvas (w, h)
    matrix: Optional[List[float]]   # PIL inverse affine matrix (6 doubles)
    fixed: Optional[Tuple[int, ...]]  # libImaging 16.16 coefficients a0..a5
    fixed_ok: bool


def _fix(v: float) -> int:
    # libImaging Geometry.c: #define FIX(v) FLOOR((v)*65536.0 + 0.5), FLOOR(x) = x >= 0 ? (int)x : (int)floor(x)
    t = v * 65536.0 + 0.5
    return int(t) if t >= 0.0 else int(math.floor(t))


def rotation_plan(w: int, h: int, angle: float) -> RotationPlan:
    """Pillow ``Image.rotate(angle, expand=True)`` parameter math (PIL/Image.py:2509-2585:
    fast paths, rounded cos/sin matrix about the centre, expanded canvas, re-centring)
    followed by libImaging ``affine_fixed`` coefficient conversion (Geometry.c).
    Called for reference transforms/rotations.py:96."""
    angle = angle % 360.0
    if angle == 0:
        return RotationPlan("copy", (w, h), None, None, True)
    if angle == 180:
        return RotationPlan("rot180", (w, h), None, None, True)
    if angle in (90, 270):
        return RotationPlan("rot90" if angle == 90 else "rot270", (h, w), None, None, True)
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0,
         round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]

    def transform(x, y, m):
        return m[0] * x + m[1] * y + m[2], m[3] * x + m[4] * y + m[5]

    m[2], m[5] = transform(-cx, -cy, m)
    m[2] += cx
    m[5] += cy
    xx, yy = [], []
    for x, y in ((0, 0), (w, 0), (w, h), (0, h)):
        tx, ty = transform(x, y, m)
        xx.append(tx)
        yy.append(ty)
    nw = math.ceil(max(xx)) - math.floor(min(xx))
    nh = math.ceil(max(yy)) - math.floor(min(yy))
    m[2], m[5] = transform(-(nw - w) / 2.0, -(nh - h) / 2.0, m)

    def in_range(x, y):
        return abs(x * m[0] + y * m[1] + m[2]) < 32768.0 and abs(x * m[3] + y * m[4] + m[5]) < 32768.0

    fixed_ok = in_range(0, 0) and in_range(nw, nh) and in_range(0, nh) and in_range(nw, 0)
    fixed = (_fix(m[0]), _fix(m[1]), _fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
             _fix(m[3]), _fix(m[4]), _fix(m[5] + m[3] * 0.5 + m[4] * 0.5))
    if m[1] == 0 and m[3] == 0:
        # libImaging would take ImagingScaleAffine here; unreachable for |sin| >= 5e-16
        fixed_ok = False
    return RotationPlan("affine", (nw, nh), m, fixed, fixed_ok)


# --------------------------------------------------------------------------- crops
def compute_crop(value, total_length) -> int:
    """reference transforms/recadrages.py:7-10."""
    if value < 0:
        raise ValueError("Les valeurs de rognage ne peuvent pas être négatives.")
    return int(total_length * value) if 0 <= value < 1 else int(value)


def border_crop_rect(width: int, height: int, crop_margins, name: str = "") -> Tuple[int, int, int, int]:
    """reference transforms/recadrages.py:37-46 -> (x0, y0, w, h)."""
    top, bottom, left, right = crop_margins
    t = compute_crop(top, height)
    b = compute_crop(bottom, height)
    l = compute_crop(left, width)
    r = compute_crop(right, width)
    if t + b >= height or l + r >= width:
        raise ValueError(f"Les marges de rognage sont trop grandes pour l'image {name}.")
    return l, t, width - l - r, height - t - b


def xywhn2xyxy_f32(boxes: np.ndarray, w: int, h: int) -> np.ndarray:
    """ultralytics.utils.ops.xywhn2xyxy as the reference calls it (crop_square.py:167):
    float64 expression, float32 result."""
    x = np.asarray(boxes)
    y = np.empty_like(x, dtype=np.float32)
    half_w, half_h = x[..., 2] / 2, x[..., 3] / 2
    y[..., 0] = w * (x[..., 0] - half_w)
    y[..., 1] = h * (x[..., 1] - half_h)
    y[..., 2] = w * (x[..., 0] + half_w)
    y[..., 3] = h * (x[..., 1] + half_h)
    return y


def square_crop_bounds(bboxes: np.ndarray, width: int, height: int):
    """reference transforms/crop_square.py:167-189: (crop_size, (lbx, ubx), (lby, uby)).
    Raises RuntimeError when no admissible position exists."""
    absolute = xywhn2xyxy_f32(bboxes, width, height)
    crop_size = min(height, width)
    x_min, y_min = absolute[:, :2].min(axis=0)
    x_max, y_max = absolute[:, 2:].max(axis=0)
    lbx = max(0, int(x_max - crop_size))
    ubx = min(int(x_min), width - crop_size)
    lby = max(0, int(y_max - crop_size))
    uby = min(int(y_min), height - crop_size)
    if lbx > ubx or lby > uby:
        raise RuntimeError(
            f"Impossible de trouver une position de crop carré valide "
            f"contenant entièrement la bbox [{x_min},{y_min},{x_max},{y_max}] "
            f"dans une image {width}x{height} avec crop_size={crop_size}. Crop annulé.")
    return crop_size, (lbx, ubx), (lby, uby)


def format_label_rows(class_ids, boxes) -> str:
    """reference transforms/crop_square.py:98-101."""
    return "".join(f"{c} {b[0]:.6f} {b[1]:.6f} {b[2]:.6f} {b[3]:.6f}\n" for c, b in zip(class_ids, boxes))


# --------------------------------------------------------------------------- overlay
def overlay_size(ov_w: int, ov_h: int, bg_w: int, bg_h: int, target_ratio: float) -> Tuple[int, int]:
    """reference transforms/overlays.py:106-126."""
    bg_diag = math.hypot(bg_w, bg_h)
    ov_diag_target = bg_diag * target_ratio
    if ov_h == 0:
        raise ValueError(f"dimensions de l'overlay invalides ({ov_w}x{ov_h}).")
    ar = ov_w / ov_h
    h_max = min(bg_w / ar, bg_h)
    max_ov_diag = math.hypot(ar * h_max, h_max)
    ov_diag = min(ov_diag_target, max_ov_diag)
    nh = int(math.sqrt(ov_diag ** 2 / (ar ** 2 + 1)))
    nw = int(ar * nh)
    return nw, nh


# --------------------------------------------------------------------------- HSV filters
def rescale_filter(filter_tuple, use_gimp_scale: bool = False):
    """reference transforms/filtres_liste.py:8-39 (same checks, same messages' intent)."""
    min_H, min_S, min_V, max_H, max_S, max_V = filter_tuple
    if not use_gimp_scale:
        if any(h > 180 for h in [min_H, max_H]):
            raise ValueError(f"Valeur(s) H ({min_H} - {max_H}) du filtre HSV au format OpenCV non conforme")
        if all(val <= 100 for val in [min_S, min_V, max_S, max_V]):
            print(f"Warning : aucune des valeurs S et V du filtre HSV au dessus de 100. ({min_S}, {min_V}, {max_S}, {max_V})."
                  "Vérifiez que votre filtre est au format OpenCV (0-180, 0-255, 0-255)"
                  "Non-bloquant, poursuite du traitement...")
        return filter_tuple
    if any(sv > 100 for sv in [min_S, min_V, max_S, max_V]):
        raise ValueError(f"Valeur(s) S et V ({min_S}, {min_V}, {max_S}, {max_V}) au delà des limites admises par GIMP."
                         "Vérifiez que votre filtre est au format de GIMP (0-360, 0-100, 0-100)")
    min_H //= 2
    max_H //= 2
    min_S *= 2.55
    min_V *= 2.55
    max_S *= 2.55
    max_V *= 2.55
    return min_H, min_S, min_V, max_H, max_S, max_V


def _round_half_even(v) -> int:
    # cv2.inRange converts scalar bounds for CV_8U images with cvRound (half to even)
    r = float(np.rint(float(v)))
    return int(max(min(r, 2 ** 30), -2 ** 30))


def pack_hsv_ranges(color_ranges, zones=None, use_gimp_scale: bool = False):
    """Validate like reference filtres_liste.py:74-81,97-107 and convert to the C-ABI
    ``ipp_hsv_range`` records (integer bounds rounded as cv2.inRange rounds them)."""
    if not color_ranges:
        raise ValueError("`color_ranges_to_exclude_hsv` est requis pour traiter les données")
    if zones and len(zones) != len(color_ranges):
        raise ValueError(f"Les zones d'application des filtres colorimétriques ({len(zones)}) ne correspondent pas "
                         f"aux filtres ({len(color_ranges)}). Les 2 paramètres doivent être de même longueur !.")
    elif not zones:
        zones = [None] * len(color_ranges)
    if len(color_ranges) > _native.IPP_MAX_HSV_RANGES:
        raise NotImplementedError(f"at most {_native.IPP_MAX_HSV_RANGES} HSV ranges per call")
    out = []
    for filt, zone in zip(color_ranges, zones):
        h0, s0, v0, h1, s1, v1 = rescale_filter(filt, use_gimp_scale)
        lo, hi = (h0, s0, v0), (h1, s1, v1)
        # cv2.inRange needs both bound arrays of one dtype: np.array([int, float, float]) is float64
        # for both or int for both; a mix of an all-int lower and a float upper raises in cv2.
        lo_f = any(isinstance(x, float) for x in lo)
        hi_f = any(isinstance(x, float) for x in hi)
        if lo_f != hi_f:
            raise ValueError("cv2.inRange: lowerb and upperb must have the same type (int vs float bounds)")
        r = _native.HsvRange()
        for b in range(3):
            r.lo[b] = _round_half_even(lo[b])
            r.hi[b] = _round_half_even(hi[b])
        z = zone if zone else (0, 0, 0, 0)
        for b in range(4):
            r.zone[b] = int(z[b])
        out.append(r)
    return out
