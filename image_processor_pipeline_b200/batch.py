"""Batched in-memory operators: what a ProcessingStep that maps ONE transform over a whole
directory computes (reference pipeline.py:304-437), as one or a few kernel launches per batch.

* ``crop_batch``      per-image crop offsets (crop_square.py:192-196)
* ``rotate_batch``    convert("RGBA") + rotate(expand=True) + getbbox per image (rotations.py:55,88-99)
* ``resize_lanczos_batch`` / ``overlay_batch``  LANCZOS resize of RGBA sprites (+ paste onto backgrounds
                      with the YOLO label) per pair (overlays.py:104-149)

Random draws come from the caller's Python generator in the reference's order (a native planner
advances the very MT19937 state).  Everything is integer and bit-exact except bilinear rotation of
opaque sources (single-precision interpolation, <= 1 LSB).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native, chain

ROT_ITEM_DTYPE = np.dtype([(n, "<i4") for n in ("src", "kind", "a0", "a1", "a2", "a3", "a4", "a5", "dst_w", "dst_h",
                                                "tile_start", "tiles_x")] + [("dst_off", "<i8"), ("m", "<f8", (6,))])
assert ROT_ITEM_DTYPE.itemsize == 104


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_batch(t: torch.Tensor, name: str, channels: Sequence[int]):
    if not isinstance(t, torch.Tensor) or t.dtype != torch.uint8 or t.dim() != 4:
        raise TypeError(f"{name} must be a torch.uint8 tensor [N, H, W, C]")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (there is no CPU fallback)")
    if t.shape[3] not in channels:
        raise ValueError(f"{name} must have {' or '.join(map(str, channels))} channels")
    return t.contiguous()


# --------------------------------------------------------------------------- crop
def crop_batch(imgs: torch.Tensor, xy, out_w: int, out_h: int) -> torch.Tensor:
    """out[i] = imgs[i, y0_i : y0_i + out_h, x0_i : x0_i + out_w] with per-image offsets ``xy`` ([N, 2] ints:
    x0, y0) — the square crop of crop_square.py:192-196 for a whole directory in one launch."""
    t = _check_batch(imgs, "imgs", (1, 2, 3, 4))
    n, H, W, c = t.shape
    xy_h = np.ascontiguousarray(np.asarray(xy.cpu() if isinstance(xy, torch.Tensor) else xy), dtype=np.int32).reshape(n, 2)
    if (xy_h < 0).any() or (xy_h[:, 0] > W - out_w).any() or (xy_h[:, 1] > H - out_h).any():
        raise ValueError("crop offsets outside the images")
    out = torch.empty((n, out_h, out_w, c), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        xy_d = chain._upload("crop_xy", xy_h, t.device)
        _native.check(_native.load().ipp_crop_batch_u8(t.data_ptr(), W * c, H * W * c, W, H, c, xy_d.data_ptr(), out_w, out_h,
                                                       out.data_ptr(), out_w * c, out_h * out_w * c, n, _stream()), "ipp_crop_batch_u8")
        xy_d.record_stream(torch.cuda.current_stream())
    return out


# --------------------------------------------------------------------------- rotation
class RotateBatchResult:
    """Canvases of a batched rotation.  ``pool``: uint8 device buffer; canvas i = ``canvas(i)`` ([dst_h, dst_w, 4]
    view, rows padded to a multiple of 4 pixels); ``bbox``: int32 [N, 4] PIL getbbox per canvas (x1 == 0: empty);
    ``fit(i)`` = the tight crop rotations.py:99-110 saves; ``angles``: the drawn angles."""

    def __init__(self, pool, items, bbox_d, angles):
        self.pool, self.items, self._bbox_d, self.angles = pool, items, bbox_d, angles
        self._bbox = None

    @property
    def bbox(self) -> np.ndarray:
        if self._bbox is None:
            self._bbox = self._bbox_d.cpu().numpy()
        return self._bbox

    def canvas(self, i: int) -> torch.Tensor:
        it = self.items[i]
        w, h = int(it["dst_w"]), int(it["dst_h"])
        wpad = (w + 3) & ~3
        off = int(it["dst_off"])
        return self.pool[off: off + wpad * 4 * h].view(h, wpad, 4)[:, :w]

    def fit(self, i: int) -> torch.Tensor:
        x0, y0, x1, y1 = (int(v) for v in self.bbox[i])
        c = self.canvas(i)
        return c if x1 <= 0 else c[y0:y1, x0:x1]


def rotate_batch(imgs: torch.Tensor, rng=_random, angle_min: float = 1.0, angle_max: float = 359.0,
                 resample: str = "nearest", angles: Optional[Sequence[float]] = None) -> RotateBatchResult:
    """One rotation per image of ``imgs`` ([N, H, W, 3] RGB — alpha 255 supplied, the convert("RGBA") of
    rotations.py:55 folded into the load — or [N, H, W, 4] RGBA): ``angle = rng.uniform(angle_min, angle_max)``
    per image in order (rotations.py:89), ``rotate(angle, expand=True)`` + ``getbbox()`` (rotations.py:96-99) in
    ONE launch.  ``resample``: "nearest" (the reference's mode, bit-exact) or "bilinear" (<= 1 LSB for RGB sources,
    exact double-precision interpolation for RGBA sources).  ``angles``: use these instead of drawing."""
    t = _check_batch(imgs, "imgs", (3, 4))
    n, H, W, c = t.shape
    lib = _native.load()
    mode = {"nearest": 0, "bilinear": 1}[resample]
    items = np.zeros(n, dtype=ROT_ITEM_DTYPE)
    ang = np.zeros(n, dtype=np.float64)
    totals = np.zeros(2, dtype=np.int64)
    if angles is None:
        bridge = chain._RngBridge(rng)
        _native.check(lib.ipp_plan_rotate_batch_host(W, H, n, float(angle_min), float(angle_max), bridge.mt.ctypes.data,
                                                     bridge.idx.ctypes.data, items.ctypes.data, ang.ctypes.data, totals.ctypes.data),
                      "ipp_plan_rotate_batch_host")
        bridge.commit()
    else:                       # fixed angles: a throw-away generator state is not needed — plan one image at a time
        if len(angles) != n:
            raise ValueError("one angle per image")
        off = tiles = 0
        one, a1, t2 = np.zeros(1, dtype=ROT_ITEM_DTYPE), np.zeros(1, dtype=np.float64), np.zeros(2, dtype=np.int64)
        mt, idx = np.zeros(624, dtype=np.uint32), np.array([624], dtype=np.int32)
        for i, a in enumerate(angles):
            _native.check(lib.ipp_plan_rotate_batch_host(W, H, 1, float(a), float(a), mt.ctypes.data, idx.ctypes.data,
                                                         one.ctypes.data, a1.ctypes.data, t2.ctypes.data), "ipp_plan_rotate_batch_host")
            items[i] = one[0]
            items[i]["src"], items[i]["tile_start"], items[i]["dst_off"] = i, tiles, off
            ang[i] = float(a)
            off += int(t2[0]); tiles += int(t2[1])
        totals[:] = (off, tiles)
    with torch.cuda.device(t.device):
        # the pool's size follows the drawn angles: rounded up to an eighth of its power of two, so that consecutive batches ask the
        # caching allocator for a size it has seen (a fresh multi-GB cudaMalloc per call cost more than the kernel)
        need = max(int(totals[0]), 256)
        gran = max(1 << 21, 1 << max(need.bit_length() - 4, 0))
        pool = torch.empty((need + gran - 1) // gran * gran, dtype=torch.uint8, device=t.device)
        bbox = torch.empty((n, 4), dtype=torch.int32, device=t.device)
        items_d = chain._upload("rot_items", items, t.device)
        _native.check(lib.ipp_rotate_batch_u8(t.data_ptr(), W * c, H * W * c, W, H, c, n, items_d.data_ptr(), n, int(totals[1]), mode,
                                              pool.data_ptr(), bbox.data_ptr(), _stream()), "ipp_rotate_batch_u8")
        items_d.record_stream(torch.cuda.current_stream())
    return RotateBatchResult(pool, items, bbox, ang)


# --------------------------------------------------------------------------- HSV key + largest component
def key_and_keep_largest(bgr: torch.Tensor, color_ranges, zones=None, use_gimp_scale: bool = False, crop_margins=(0, 0, 0, 0)):
    """``process_images_with_color_masks`` then ``keep_largest_component`` (filtres_liste.py:90-134, pixels_isolés.py:29-58)
    for a batch [N, H, W, 3] of BGR images WITHOUT the BGRA image in between: the frame is read once into a 1-bit mask, the
    largest 8-connected component is found on bits, and the BGRA result is written once.  Returns (bgra [N, h, w, 4], info
    int32 CPU [N, 8] = {x, y, w, h of the kept component, area, n_fg, 0, 0}); ``_crop_fit`` of the reference is the view
    bgra[i, y:y+h, x:x+w] (w == 0: nothing left, the reference writes no file)."""
    from . import hostmath
    t = _check_batch(bgr, "bgr", (3,))
    n, FH, FW, _ = t.shape
    lib = _native.load()
    rect = hostmath.border_crop_rect(FW, FH, crop_margins)
    ranges = hostmath.pack_hsv_ranges(color_ranges, zones, use_gimp_scale)
    geom = chain._geom_struct(FW, FH, rect, rect[2], rect[3], ranges)
    w, h = rect[2], rect[3]
    out = torch.empty((n, h, w, 4), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        ws_bytes = lib.ipp_chain_mask_workspace_bytes(C.byref(geom), n)
        ws = chain._scratch("kk_mask", ws_bytes, t.device)
        info = torch.empty((n, 8), dtype=torch.int32, device=t.device)
        _native.check(lib.ipp_chain_stage_a(t.data_ptr(), FW * 3, FH * FW * 3, C.byref(geom), info.data_ptr(), ws.data_ptr(), ws_bytes, n,
                                            _stream()), "ipp_chain_stage_a")
        _native.check(lib.ipp_bgra_from_keep_u8(t.data_ptr(), FW * 3, FH * FW * 3, C.byref(geom), ws.data_ptr(), out.data_ptr(), w * 4, h * w * 4, n,
                                                _stream()), "ipp_bgra_from_keep_u8")
        info_h = info.cpu().numpy()
    return out, info_h


# --------------------------------------------------------------------------- photometric jitter (tranfo.py)
ENH_ITEM_DTYPE = np.dtype([("brightness", "<f4"), ("contrast", "<f4"), ("color", "<f4"), ("blur", "<i4"), ("box_radius", "<i4"),
                           ("ww", "<u4"), ("fw", "<u4"), ("lut", "<i4")])
assert ENH_ITEM_DTYPE.itemsize == 32


def sample_flags(n: int, rng=_random):
    """The 'sample' pairing of pipeline.py:209-225: two ``random.sample(files, int(n * 0.3))`` draws decide which of the n
    files get the blur / the per-channel jitter.  Returns (do_blur, do_rgb) as bool arrays."""
    files = list(range(n))
    blur = set(rng.sample(files, int(n * 0.3)))
    do_blur = np.array([i in blur for i in files], dtype=bool)
    rgb = set(rng.sample(files, int(n * 0.3)))
    do_rgb = np.array([i in rgb for i in files], dtype=bool)
    return do_blur, do_rgb


def plan_enhance(n: int, do_blur, do_rgb, rng=_random):
    """Per-image draws of tranfo.py:38-51 in order (native planner on the generator's own state).
    Returns (items, luts [n, 3, 256] uint8, draws [n, 4] float64)."""
    do_blur = np.ascontiguousarray(np.asarray(do_blur, dtype=bool).astype(np.uint8))
    do_rgb = np.ascontiguousarray(np.asarray(do_rgb, dtype=bool).astype(np.uint8))
    if len(do_blur) != n or len(do_rgb) != n:
        raise ValueError("one blur / rgb flag per image")
    items = np.zeros(n, dtype=ENH_ITEM_DTYPE)
    luts = np.zeros((n, 3, 256), dtype=np.uint8)
    draws = np.zeros((n, 4), dtype=np.float64)
    bridge = chain._RngBridge(rng)
    _native.check(_native.load().ipp_plan_enhance_host(n, do_blur.ctypes.data, do_rgb.ctypes.data, bridge.mt.ctypes.data, bridge.idx.ctypes.data,
                                                       items.ctypes.data, luts.ctypes.data, draws.ctypes.data), "ipp_plan_enhance_host")
    bridge.commit()
    return items, luts, draws


def enhance_batch(images: torch.Tensor, do_blur, do_rgb, rng=_random, inplace: bool = False, plan=None) -> torch.Tensor:
    """``enhance_image`` (tranfo.py:37-51) for a batch [N, H, W, 3] of RGB images: Brightness, Contrast, Color jitter, GaussianBlur
    where ``do_blur[i]``, per-channel 256-entry tables where ``do_rgb[i]`` — draws per image in the reference's order from ``rng``.
    Bit-exact against Pillow.  ``plan``: (items, luts) from ``plan_enhance`` (then nothing is drawn here)."""
    t = _check_batch(images, "images", (3,))
    n, H, W, _ = t.shape
    items, luts = plan_enhance(n, do_blur, do_rgb, rng)[:2] if plan is None else plan
    lib = _native.load()
    out = t if inplace else torch.empty_like(t)
    blur_idx = np.flatnonzero(items["blur"]).astype(np.int32)
    with torch.cuda.device(t.device):
        items_d = chain._upload("enh_items", items, t.device)
        luts_d = chain._upload("enh_luts", luts, t.device)
        blur_d = chain._upload("enh_blur", blur_idx if len(blur_idx) else np.zeros(1, np.int32), t.device)
        ws_bytes = lib.ipp_enhance_workspace_bytes(W, H, n, len(blur_idx))
        ws = chain._scratch("enh_ws", ws_bytes, t.device)
        _native.check(lib.ipp_enhance_batch_u8(t.data_ptr(), out.data_ptr(), W, H, n, items_d.data_ptr(), luts_d.data_ptr(), blur_d.data_ptr(),
                                               len(blur_idx), ws.data_ptr(), ws_bytes, _stream()), "ipp_enhance_batch_u8")
        for x in (items_d, luts_d, blur_d):
            x.record_stream(torch.cuda.current_stream())
    return out


# --------------------------------------------------------------------------- LANCZOS resize / overlay
def _resize_items(items: np.ndarray, sprites: torch.Tensor, device_tables: bool, eps: float):
    """Layout + tables + premultiplied planes + tensor-core passes for an item array over ``sprites``
    ([S, h, w, 4] RGBA).  Returns (totals, items_d, spr_ws): resized RGBA sprites at items[i]['spr_off']."""
    lib = _native.load()
    dev = sprites.device
    S, sh, sw, _ = sprites.shape
    tot_s = _native.PlanTotals()
    _native.check(lib.ipp_chain_layout_planar_host(items.ctypes.data, len(items), C.byref(tot_s)), "ipp_chain_layout_planar_host")
    tot = dict(table_ints=int(tot_s.table_ints), max_rows=int(tot_s.max_rows), ws=int(tot_s.sprite_ws_bytes), ht=int(tot_s.h_tiles),
               vt=int(tot_s.v_tiles), band_bytes=int(tot_s.band_units) * 8, max_band_warps=int(tot_s.max_band_warps),
               vstaged_smem=int(tot_s.vstaged_smem_bytes))
    K = len(items)
    items_d = chain._upload("ov_items", items, dev)
    spr_ws = chain._scratch("ov_sprites", max(tot["ws"], 256), dev)
    band_d = chain._scratch("ov_bands", max(tot["band_bytes"], 256), dev)
    if device_tables:
        tab_d = chain._scratch("ov_tables", max(1, tot["table_ints"]) * 4, dev, torch.int32)
        amb_d = chain._scratch("ov_amb", chain.amb_ints() * 4, dev, torch.int32)
        _native.check(lib.ipp_chain_tables(items_d.data_ptr(), K, tot["max_rows"], tab_d.data_ptr(), amb_d.data_ptr(),
                                           chain.AMB_CAPACITY, float(eps), _stream()), "ipp_chain_tables")
        amb_host = chain._pinned("ov_amb", chain.AMB_REC * (chain.AMB_INLINE + 1) * 4, torch.int32)
        amb_host.copy_(amb_d[:chain.AMB_REC * (chain.AMB_INLINE + 1)], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
    else:
        tab_d = torch.from_numpy(chain._host_tables(items, lib, tot["table_ints"])).to(dev)
    # the planes do not depend on the tables: queued before the host looks at the flagged rows
    _native.check(lib.ipp_sprite_planes_u8(sprites.data_ptr(), sw * 4, sh * sw * 4, sw, sh, S, items_d.data_ptr(), K, spr_ws.data_ptr(),
                                           _stream()), "ipp_sprite_planes_u8")
    if device_tables:
        ev.synchronize()
        n_amb = int(amb_host[0])
        if n_amb:
            chain._confirm_rows(lib, items, amb_d, amb_host, n_amb, tab_d, dev)
    geom = chain._geom_struct(sw, sh, (0, 0, sw, sh), sw, sh, [_native.HsvRange()])
    _native.check(lib.ipp_chain_stage_b2(sprites.data_ptr(), sw * 4, sh * sw * 4, C.byref(geom), items_d.data_ptr(), K, 0, tot["ht"], tot["vt"],
                                         tab_d.data_ptr(), band_d.data_ptr(), tot["max_band_warps"], 0, tot["vstaged_smem"],
                                         spr_ws.data_ptr(), spr_ws.data_ptr(), _stream()), "ipp_chain_stage_b2")
    # size unchanged: Image.resize returns a plain copy (PIL/Image.py: `if self.size == size: return self.copy()`), i.e. NO
    # RGBA -> RGBa -> RGBA round trip, which is lossy for 0 < alpha < 255
    for i in np.flatnonzero((items["new_w"] == items["spr_w"]) & (items["new_h"] == items["spr_h"])):
        off, nbytes = int(items[i]["spr_off"]), sw * sh * 4
        spr_ws[off: off + nbytes].copy_(sprites[int(items[i]["frame"])].reshape(-1))
    return tot, items_d, spr_ws


def resize_lanczos_batch(sprites: torch.Tensor, sizes: Sequence[Tuple[int, int]], which: Optional[Sequence[int]] = None,
                         device_tables: Optional[bool] = None, ambiguity_eps: float = chain.AMBIGUITY_EPS) -> List[torch.Tensor]:
    """``Image.resize(size, LANCZOS)`` (overlays.py:129) of RGBA sprites on the tensor cores (exact integer banded
    GEMMs, bit-exact): output i = sprites[which[i]] (default i) resized to sizes[i] = (new_w, new_h)."""
    t = _check_batch(sprites, "sprites", (4,))
    S, sh, sw, _ = t.shape
    K = len(sizes)
    which = list(range(K)) if which is None else list(which)
    items = np.zeros(K, dtype=chain.ITEM_DTYPE)
    for i, ((nw, nh), s) in enumerate(zip(sizes, which)):
        if nw <= 0 or nh <= 0:
            raise ValueError("height and width must be > 0")          # PIL's message for empty sizes
        if not 0 <= s < S:
            raise ValueError("sprite index out of range")
        items[i]["frame"], items[i]["kind"] = s, chain.KIND["copy"]
        items[i]["cc_w"], items[i]["cc_h"], items[i]["spr_w"], items[i]["spr_h"] = sw, sh, sw, sh
        items[i]["new_w"], items[i]["new_h"] = int(nw), int(nh)
    if K == 0:
        return []
    if device_tables is None:
        device_tables = K >= 8                      # few sprites: exact host tables are cheaper than the device + confirm round trip
    with torch.cuda.device(t.device):
        tot, items_d, spr_ws = _resize_items(items, t, device_tables, ambiguity_eps)
        offs = np.zeros(K, dtype=np.int64)
        total = 0
        for i in range(K):
            offs[i] = total
            total += int(items[i]["new_w"]) * int(items[i]["new_h"]) * 4
        out = torch.empty(total, dtype=torch.uint8, device=t.device)
        _native.check(_native.load().ipp_sprites_export_u8(items.ctypes.data, K, spr_ws.data_ptr(), out.data_ptr(), offs.ctypes.data, _stream()),
                      "ipp_sprites_export_u8")
        items_d.record_stream(torch.cuda.current_stream())
    return [out[int(offs[i]): int(offs[i]) + int(items[i]["new_w"]) * int(items[i]["new_h"]) * 4].view(int(items[i]["new_h"]), int(items[i]["new_w"]), 4)
            for i in range(K)]


class OverlayBatchResult:
    """images: [K, H, W, 3] composites; labels: YOLO label text per image (overlays.py:149); pair: index of the
    (sprite, background) pair each image came from (pairs whose resize fails are absent); items: plan records."""

    def __init__(self, images, labels, items):
        self.images, self.labels, self.items = images, labels, items
        self.pair = items["cc_x"].copy()


def overlay_batch(sprites: torch.Tensor, backgrounds: torch.Tensor, n_pairs: Optional[int] = None, rng=_random, yolo_class_id: int = 0,
                  scale_min: float = 0.15, scale_max: float = 0.30, ambiguity_eps: float = chain.AMBIGUITY_EPS) -> OverlayBatchResult:
    """``paste_overlay_onto_background`` (overlays.py:104-149) for ``n_pairs`` pairs (sprite k % S, background k % M):
    per pair ``uniform`` (scale), LANCZOS resize, ``randint`` x 2 (position), alpha paste, label — draws in the
    reference's order from ``rng``.  sprites: [S, h, w, 4] RGBA, backgrounds: [M, H, W, 3] RGB (CUDA)."""
    s = _check_batch(sprites, "sprites", (4,))
    b = _check_batch(backgrounds, "backgrounds", (3,))
    if s.device != b.device:
        raise ValueError("sprites and backgrounds must be on the same device")
    S, sh, sw, _ = s.shape
    M, BH, BW, _ = b.shape
    n_pairs = max(S, M) if n_pairs is None else int(n_pairs)
    lib = _native.load()
    pp = _native.PlanParams()
    pp.yolo_class_id, pp.scale_min, pp.scale_max = int(yolo_class_id), float(scale_min), float(scale_max)
    pp.bg_w, pp.bg_h, pp.n_bgs = BW, BH, M
    items = np.zeros(max(1, n_pairs), dtype=chain.ITEM_DTYPE)
    labels = np.zeros(max(1, n_pairs) * 64, dtype=np.uint8)
    n_items = C.c_int32(0)
    totals = _native.PlanTotals()
    bridge = chain._RngBridge(rng)
    _native.check(lib.ipp_plan_overlay_batch_host(sw, sh, S, n_pairs, C.byref(pp), bridge.mt.ctypes.data, bridge.idx.ctypes.data,
                                                  items.ctypes.data, C.addressof(n_items), labels.ctypes.data, labels.size, C.byref(totals)),
                  "ipp_plan_overlay_batch_host")
    bridge.commit()
    K = n_items.value
    items = items[:K].copy()
    text = labels[:totals.labels_bytes].tobytes().decode("ascii")
    lab = text.split("\0")[:K] if K else []
    out = torch.empty((K, BH, BW, 3), dtype=torch.uint8, device=b.device)
    if K:
        with torch.cuda.device(b.device):
            tot, items_d, spr_ws = _resize_items(items, s, True, ambiguity_eps)
            _native.check(lib.ipp_chain_stage_c_rect(b.data_ptr(), BW * 3, BH * BW * 3, BW, BH, items_d.data_ptr(), K, spr_ws.data_ptr(),
                                                     out.data_ptr(), BW * 3, BH * BW * 3, BW, BH, _stream()), "ipp_chain_stage_c_rect")
            items_d.record_stream(torch.cuda.current_stream())
    return OverlayBatchResult(out, lab, items)
