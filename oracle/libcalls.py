"""T2 oracle: the reference's own cv2 / Pillow / numpy call sequences, in memory.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Every function issues the
same library calls, in the same order and with the same arguments, as the
reference transform it cites; only file decode/encode is removed (arrays in,
arrays out).  Random draws use the module-level ``random`` generator in the
reference's draw order unless explicit parameters are given.

Array conventions: cv2-side functions take/return HWC uint8 in BGR(A) order,
Pillow-side functions take/return HWC uint8 in RGB(A) order — exactly what the
reference's ``cv2.imread`` / ``Image.open`` would hand to the arithmetic.
"""
from __future__ import annotations

import math
import random
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np
from PIL import Image

ALL_SYMS = ("o", "h", "v", "hv")


# --------------------------------------------------------------------------- symmetry
def symmetries(image: np.ndarray, syms: Sequence[str]) -> List[np.ndarray]:
    """reference transforms/symmetry.py:114-131 — copy / cv2.flip(1|0|-1)."""
    gen = {
        "o": lambda img: img.copy(),
        "h": lambda img: cv2.flip(img, 1),
        "v": lambda img: cv2.flip(img, 0),
        "hv": lambda img: cv2.flip(img, -1),
    }
    return [gen[s](image) for s in syms]


def select_symmetries(pool=None, choose_random=None, include_original=True) -> List[str]:
    """reference transforms/symmetry.py:97-126 — which variants, in which order."""
    pool = pool if pool else list(ALL_SYMS)
    choose_random = len(pool) if choose_random is None else choose_random
    filter_ = random.sample(pool, choose_random)
    if include_original and "o" not in set(filter_):
        filter_.append("o")
    return filter_


# --------------------------------------------------------------------------- rotation
def rotate_expand(rgba: np.ndarray, angle: float, resample=Image.Resampling.NEAREST) -> np.ndarray:
    """reference transforms/rotations.py:96 — ``img.rotate(angle, expand=True)``."""
    img = Image.fromarray(rgba, "RGBA")
    return np.asarray(img.rotate(angle, resample=resample, expand=True))


def rotate_and_fit(rgba: np.ndarray, angle: float, resample=Image.Resampling.NEAREST):
    """reference transforms/rotations.py:96-110 — rotate, getbbox, crop.

    Returns (image, bbox) with bbox=None when getbbox() found nothing (the
    reference then keeps the un-cropped canvas)."""
    img = Image.fromarray(rgba, "RGBA")
    rotated = img.rotate(angle, resample=resample, expand=True)
    bbox = rotated.getbbox()
    if bbox:
        cropped = rotated.crop(bbox)
        if cropped.width > 0 and cropped.height > 0:
            return np.asarray(cropped), bbox
    return np.asarray(rotated), None


def to_rgba(arr: np.ndarray) -> np.ndarray:
    """reference transforms/rotations.py:55 — ``Image.open(p).convert('RGBA')`` for
    an already-decoded L / RGB / RGBA array."""
    if arr.ndim == 2:
        return np.asarray(Image.fromarray(arr, "L").convert("RGBA"))
    if arr.shape[2] == 3:
        return np.asarray(Image.fromarray(arr, "RGB").convert("RGBA"))
    return np.asarray(Image.fromarray(arr, "RGBA").convert("RGBA"))


# --------------------------------------------------------------------------- crops
def compute_crop(value, total_length) -> int:
    """reference transforms/recadrages.py:7-10."""
    if value < 0:
        raise ValueError("Les valeurs de rognage ne peuvent pas être négatives.")
    return int(total_length * value) if 0 <= value < 1 else int(value)


def crop_from_border(image: np.ndarray, crop_margins=(0, 0, 0, 0)) -> np.ndarray:
    """reference transforms/recadrages.py:34-46."""
    top, bottom, left, right = crop_margins
    height, width = image.shape[:2]
    t = compute_crop(top, height)
    b = compute_crop(bottom, height)
    l = compute_crop(left, width)
    r = compute_crop(right, width)
    if t + b >= height or l + r >= width:
        raise ValueError("Les marges de rognage sont trop grandes pour l'image.")
    return image[t:height - b, l:width - r]


def fit_crop(arr: np.ndarray) -> np.ndarray:
    """reference transforms/recadrages.py:71-77 — PIL getbbox + crop (copy if empty)."""
    mode = {1: "L", 2: "LA", 3: "RGB", 4: "RGBA"}[1 if arr.ndim == 2 else arr.shape[2]]
    img = Image.fromarray(arr if arr.ndim == 3 else arr, mode)
    bbox = img.getbbox()
    out = img.copy() if not bbox else img.crop(bbox)
    return np.asarray(out)


def xywhn2xyxy(x, w=640, h=640, padw=0, padh=0):
    """ultralytics.utils.ops.xywhn2xyxy (un-vendored dependency; restated from the
    published source, float32 result — SURVEY.md §8c)."""
    x = np.asarray(x)
    y = np.empty_like(x, dtype=np.float32)
    xc, yc, xw, xh = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    half_w, half_h = xw / 2, xh / 2
    y[..., 0] = w * (xc - half_w) + padw
    y[..., 1] = h * (yc - half_h) + padh
    y[..., 2] = w * (xc + half_w) + padw
    y[..., 3] = h * (yc + half_h) + padh
    return y


def xyxy2xywhn(x, w=640, h=640, clip=False, eps=0.0):
    """ultralytics.utils.ops.xyxy2xywhn (restated, float32 result)."""
    x = np.asarray(x)
    y = np.empty_like(x, dtype=np.float32)
    x1, y1, x2, y2 = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    y[..., 0] = ((x1 + x2) / 2) / w
    y[..., 1] = ((y1 + y2) / 2) / h
    y[..., 2] = (x2 - x1) / w
    y[..., 3] = (y2 - y1) / h
    return y


def square_crop_bounds(bboxes: np.ndarray, width: int, height: int):
    """reference transforms/crop_square.py:167-189 — admissible (x0, y0) ranges."""
    bboxes_absolute = xywhn2xyxy(bboxes, width, height)
    crop_size = min(height, width)
    x_min, y_min = bboxes_absolute[:, :2].min(axis=0)
    x_max, y_max = bboxes_absolute[:, 2:].max(axis=0)
    lbx = max(0, int(x_max - crop_size))
    ubx = min(int(x_min), width - crop_size)
    lby = max(0, int(y_max - crop_size))
    uby = min(int(y_min), height - crop_size)
    if lbx > ubx or lby > uby:
        raise RuntimeError("Impossible de trouver une position de crop carré valide")
    return bboxes_absolute, crop_size, (lbx, ubx), (lby, uby)


def square_crop(image: np.ndarray, class_ids: np.ndarray, bboxes: np.ndarray,
                x0: Optional[int] = None, y0: Optional[int] = None):
    """reference transforms/crop_square.py:163-217 with the three one-token repairs
    of SURVEY.md §8a-4 (``&`` instead of ``and`` on arrays is the only one that
    touches this in-memory part).  Draws x0, y0 with random.randint like
    crop_square.py:192-193 unless given."""
    height, width = image.shape[:2]
    bboxes_absolute, crop_size, (lbx, ubx), (lby, uby) = square_crop_bounds(bboxes, width, height)
    if x0 is None:
        x0 = random.randint(lbx, ubx)
        y0 = random.randint(lby, uby)
    cropped = image[y0:y0 + crop_size, x0:x0 + crop_size]
    if cropped.size == 0:
        raise RuntimeError("Le crop a produit une image vide.")
    shifted = bboxes_absolute - np.array([[x0, y0, x0, y0]])
    clipped = np.zeros_like(shifted)
    clipped[:, 0] = np.clip(shifted[:, 0], 0, crop_size)
    clipped[:, 1] = np.clip(shifted[:, 1], 0, crop_size)
    clipped[:, 2] = np.clip(shifted[:, 2], 0, crop_size)
    clipped[:, 3] = np.clip(shifted[:, 3], 0, crop_size)
    valid = (clipped[:, 0] < clipped[:, 2]) & (clipped[:, 1] < clipped[:, 3])
    if not any(valid):
        raise RuntimeError("Aucune bbox résiduelle après le crop.")
    new_abs = clipped[valid]
    new_cls = class_ids[valid]
    new_boxes = xyxy2xywhn(new_abs, crop_size, crop_size)
    return cropped, new_cls, new_boxes, (x0, y0)


def format_label_rows(class_ids, boxes, trailing_newline=True) -> str:
    """reference transforms/crop_square.py:98-101 (one row per box, '\\n' each)."""
    rows = []
    for cls_id, box in zip(class_ids, boxes):
        cx, cy, w, h = box
        rows.append(f"{cls_id} {cx:.6f} {cy:.6f} {w:.6f} {h:.6f}\n")
    return "".join(rows)


# --------------------------------------------------------------------------- overlay
def overlay_size(ov_w: int, ov_h: int, bg_w: int, bg_h: int, target_ratio: float) -> Tuple[int, int]:
    """reference transforms/overlays.py:106-126 — float64 size math."""
    bg_diag = math.hypot(bg_w, bg_h)
    ov_diag_target = bg_diag * target_ratio
    if ov_h == 0:
        raise ValueError("dimensions de l'overlay invalides")
    ar = ov_w / ov_h
    h_max = min(bg_w / ar, bg_h)
    max_ov_diag = math.hypot(ar * h_max, h_max)
    ov_diag = min(ov_diag_target, max_ov_diag)
    nh = int(math.sqrt(ov_diag ** 2 / (ar ** 2 + 1)))
    nw = int(ar * nh)
    return nw, nh


def resize_lanczos(rgba: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """reference transforms/overlays.py:129 — ``overlay.resize((w, h), LANCZOS)``."""
    return np.asarray(Image.fromarray(rgba, "RGBA").resize(size, Image.Resampling.LANCZOS))


def paste(bg_rgb: np.ndarray, sprite_rgba: np.ndarray, pos: Tuple[int, int]) -> np.ndarray:
    """reference transforms/overlays.py:138-139 — ``bg.copy().paste(spr, pos, spr)``."""
    composite = Image.fromarray(bg_rgb, "RGB").copy()
    spr = Image.fromarray(sprite_rgba, "RGBA")
    composite.paste(spr, pos, spr)
    return np.asarray(composite)


def overlay_onto_background(sprite_rgba: np.ndarray, bg_rgb: np.ndarray, yolo_class_id=0,
                            scale_min=0.15, scale_max=0.30, params=None):
    """reference transforms/overlays.py:104-149.  ``params`` = (ratio, pos_x, pos_y)
    overrides the three random draws (uniform, randint, randint)."""
    bg_h, bg_w = bg_rgb.shape[:2]
    ov_h, ov_w = sprite_rgba.shape[:2]
    ratio = random.uniform(scale_min, scale_max) if params is None else params[0]
    nw, nh = overlay_size(ov_w, ov_h, bg_w, bg_h, ratio)
    resized = resize_lanczos(sprite_rgba, (nw, nh))
    if params is None:
        pos_x = random.randint(0, bg_w - nw)
        pos_y = random.randint(0, bg_h - nh)
    else:
        pos_x, pos_y = params[1], params[2]
    composite = paste(bg_rgb, resized, (pos_x, pos_y))
    bbox_xyxy_abs = np.array([pos_x, pos_y, pos_x + nw, pos_y + nh]).reshape(1, 4)
    cx, cy, w_norm, h_norm = xyxy2xywhn(bbox_xyxy_abs, bg_w, bg_h)[0]
    label = f"{yolo_class_id} {cx:.6f} {cy:.6f} {w_norm:.6f} {h_norm:.6f}"
    return composite, label, (nw, nh, pos_x, pos_y)


# --------------------------------------------------------------------------- HSV key
def rescale_filter(filter_tuple, use_gimp_scale=False):
    """reference transforms/filtres_liste.py:8-39."""
    min_H, min_S, min_V, max_H, max_S, max_V = filter_tuple
    if not use_gimp_scale:
        if any(h > 180 for h in [min_H, max_H]):
            raise ValueError("Valeur(s) H du filtre HSV au format OpenCV non conforme")
        return filter_tuple
    if any(sv > 100 for sv in [min_S, min_V, max_S, max_V]):
        raise ValueError("Valeur(s) S et V au delà des limites admises par GIMP.")
    min_H //= 2
    max_H //= 2
    min_S *= 2.55
    min_V *= 2.55
    max_S *= 2.55
    max_V *= 2.55
    return min_H, min_S, min_V, max_H, max_S, max_V


def color_masks_to_alpha(image_bgr: np.ndarray, color_ranges, zones=None, use_gimp_scale=False) -> np.ndarray:
    """reference transforms/filtres_liste.py:74-134 (the discarded first inRange
    of :110 is dropped)."""
    if not color_ranges:
        raise ValueError("`color_ranges_to_exclude_hsv` est requis")
    if zones and len(zones) != len(color_ranges):
        raise ValueError("zones / filtres de longueurs différentes")
    elif not zones:
        zones = [None] * len(color_ranges)
    hsv = cv2.cvtColor(image_bgr, cv2.COLOR_BGR2HSV)
    combined = np.zeros((hsv.shape[0], hsv.shape[1]), dtype=np.uint8)
    for filt, zone in zip(color_ranges, zones):
        h_min, s_min, v_min, h_max, s_max, v_max = rescale_filter(filt, use_gimp_scale)
        z_top, z_bottom, z_left, z_right = zone if zone else (0, 0, 0, 0)
        zone_mask = np.zeros_like(hsv[..., 0], dtype=np.uint8)
        zone_mask[z_top:hsv.shape[0] - z_bottom, z_left:hsv.shape[1] - z_right] = 255
        lower = np.array([h_min, s_min, v_min])
        upper = np.array([h_max, s_max, v_max])
        full_mask = cv2.inRange(hsv, lower, upper)
        current = cv2.bitwise_and(full_mask, zone_mask)
        combined = cv2.bitwise_or(combined, current)
    keep = cv2.bitwise_not(combined)
    b, g, r = cv2.split(image_bgr)
    return cv2.merge((b, g, r, keep))


# --------------------------------------------------------------------------- largest component
def crop_fit_alpha(img_bgra: np.ndarray) -> np.ndarray:
    """reference transforms/pixels_isolés.py:74-81."""
    non_transparents = cv2.findNonZero(img_bgra[:, :, 3])
    x, y, w, h = cv2.boundingRect(non_transparents)
    return img_bgra[y:y + h, x:x + w]


def keep_largest_component(image_bgra: np.ndarray, min_component_size: int = 500, with_info=False):
    """reference transforms/pixels_isolés.py:29-61.  The per-label Python loops of
    :40-52 are replaced by equivalent vectorised numpy (argmax with first-wins
    tie-break; the min_component_size loop is a no-op) when
    ``fast`` timing is wanted; results are identical."""
    if image_bgra.shape[2] != 4:
        raise AttributeError("pas de canal alpha")
    b, g, r, alpha = cv2.split(image_bgra)
    _, binary_alpha = cv2.threshold(alpha, 1, 255, cv2.THRESH_BINARY)
    num_labels, labels, stats, _ = cv2.connectedComponentsWithStats(binary_alpha, connectivity=8)
    max_area = 0
    max_label = 0
    for i in range(1, num_labels):
        area = stats[i, cv2.CC_STAT_AREA]
        if area > max_area:
            max_area = area
            max_label = i
    clean_mask = np.where(labels == max_label, 255, 0).astype(np.uint8)
    # reference :50-52 — dead loop (those pixels are already 0); kept out for speed.
    alpha = alpha.copy()
    alpha[clean_mask == 0] = 0
    cleaned = cv2.merge((b, g, r, alpha))
    cropped = crop_fit_alpha(cleaned)
    if with_info:
        nz = cv2.findNonZero(cleaned[:, :, 3])
        bbox = cv2.boundingRect(nz)
        return cropped, cleaned, bbox, int(max_area)
    return cropped


# --------------------------------------------------------------------------- photometric jitter (tranfo.py)
def enhance_image(rgb: np.ndarray, apply_blur: bool, apply_rgb: bool) -> np.ndarray:
    """reference transforms/tranfo.py:37-51 — the identical Pillow calls on an RGB array, draws from ``random``:
    Brightness / Contrast / Color enhance(uniform(.7, 1.3)), optional GaussianBlur(uniform(.5, 3)), optional
    per-channel ``point(lambda p: max(0, min(255, p * uniform(.75, 1.25))))`` (256 draws per channel)."""
    from PIL import ImageEnhance, ImageFilter
    img = Image.fromarray(rgb, "RGB")
    img = ImageEnhance.Brightness(img).enhance(random.uniform(0.7, 1.3))
    img = ImageEnhance.Contrast(img).enhance(random.uniform(0.7, 1.3))
    img = ImageEnhance.Color(img).enhance(random.uniform(0.7, 1.3))
    if apply_blur:
        img = img.filter(ImageFilter.GaussianBlur(random.uniform(0.5, 3)))
    if apply_rgb:
        r, g, b = img.split()
        r = r.point(lambda p: max(0, min(255, p * random.uniform(0.75, 1.25))))
        g = g.point(lambda p: max(0, min(255, p * random.uniform(0.75, 1.25))))
        b = b.point(lambda p: max(0, min(255, p * random.uniform(0.75, 1.25))))
        img = Image.merge("RGB", (r, g, b))
    return np.asarray(img)


def sample_pairing_flags(n: int):
    """reference pipeline.py:209-225 ('sample' pairing): which of n files get the blur / the RGB jitter — two
    ``random.sample(files, int(n * 0.3))`` draws."""
    files = list(range(n))
    blur = set(random.sample(files, int(n * 0.3)))
    do_blur = [i in blur for i in files]
    rgb = set(random.sample(files, int(n * 0.3)))
    do_rgb = [i in rgb for i in files]
    return do_blur, do_rgb
