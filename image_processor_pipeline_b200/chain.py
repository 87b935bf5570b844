"""The fused, batched, device-resident transform chain (host side).

Reproduces, for a batch of decoded frames and backgrounds already on the GPU, what the
reference's pipeline computes when its steps are chained in their natural order
(SURVEY.md §7.2):

    crop_from_border -> process_images_with_color_masks -> keep_largest_component
    -> process_rotations -> generate_symmetries -> paste_overlay_onto_background
    -> process_square_crop_around_bbox

Semantics kept from the reference:
* steps run one after the other over ALL items (pipeline.py:564-566), so random draws
  happen step-major, in item order, from Python's ``random``;
* a step lists its input directory SORTED by file name (pipeline.py:137): rotation
  outputs ``_r000 < _r001 < ...`` and symmetry outputs ``_h < _hv < _o < _v``;
* an item whose step fails or returns nothing simply does not exist for later steps;
* overlays pair item k with background k % len(backgrounds) ('modulo' pairing,
  pipeline.py:204-207; the shuffle of :202 is the caller's business).

Device work is three stages separated only where a data-dependent size has to reach
the host RNG (randint bounds depend on the sprite size): stage A (mask + largest
component), stage B1 (rotated bbox), stage B2 + C (sprite gather, LANCZOS, paste + crop).
"""
from __future__ import annotations

import ctypes as C
import random as _random
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native, hostmath

SYM_ORDER = {"h": 0, "hv": 1, "o": 2, "v": 3}        # file-name order of `<stem>_<sym>.png`
KIND = {"affine": 0, "copy": 1, "rot180": 2, "rot90": 3, "rot270": 4}

ROT_DTYPE = np.dtype([(n, "<i4") for n in
                      ("frame", "cc_x", "cc_y", "cc_w", "cc_h", "kind", "a0", "a1", "a2", "a3", "a4", "a5",
                       "canvas_w", "canvas_h", "tile_start", "_pad")])
ITEM_DTYPE = np.dtype([(n, "<i4") for n in
                       ("frame", "bg", "cc_x", "cc_y", "cc_w", "cc_h", "kind", "a0", "a1", "a2", "a3", "a4", "a5",
                        "fit_x0", "fit_y0", "spr_w", "spr_h", "flip_h", "flip_v", "new_w", "new_h", "pos_x", "pos_y",
                        "sq_x0", "sq_y0", "sq_size", "h_ksize", "v_ksize", "hb_off", "hk_off", "vb_off", "vk_off",
                        "gather_tile_start", "h_tile_start", "v_tile_start",
                        "_p0", "_p1", "_p2", "_p3", "_p4")] +
                      [("rot_off", "<i8"), ("tmp_off", "<i8"), ("spr_off", "<i8")])
assert ROT_DTYPE.itemsize == 64 and ITEM_DTYPE.itemsize == 184


@dataclass
class ChainParams:
    # crop_from_border (recadrages.py:13-18)
    crop_margins: Tuple[float, float, float, float] = (0, 0, 0, 0)
    # process_images_with_color_masks (filtres_liste.py:41-49)
    color_ranges_to_exclude_hsv: Sequence = ()
    zones: Optional[Sequence] = None
    use_gimp_scale: bool = False
    # process_rotations (rotations.py:6-19)
    num_rotations: int = 10
    include_original: bool = True
    angle_min: float = 1.0
    angle_max: float = 359.0
    # generate_symmetries (symmetry.py:11-20)
    pool: Optional[List[str]] = None
    choose_random: Optional[int] = None
    sym_include_original: bool = True
    # paste_overlay_onto_background (overlays.py:24-34)
    yolo_class_id: int = 0
    scale_min: float = 0.15
    scale_max: float = 0.30


@dataclass
class ChainResult:
    images: torch.Tensor                 # [K, cs, cs, 3] uint8 RGB on the device
    labels: List[str]                    # YOLO label file content per image (crop_square.py:98-101)
    records: List[dict] = field(default_factory=list)   # provenance / parameters per image


def _geom_struct(frame_w, frame_h, rect, bg_w, bg_h, ranges):
    g = _native.ChainGeom()
    g.frame_w, g.frame_h = frame_w, frame_h
    g.crop_l, g.crop_t, g.crop_w, g.crop_h = rect
    g.bg_w, g.bg_h = bg_w, bg_h
    g.n_ranges = len(ranges)
    for i, r in enumerate(ranges):
        g.ranges[i] = r
    return g


def _stream():
    return torch.cuda.current_stream().cuda_stream


class ChainPlanError(RuntimeError):
    pass


def plan_rotations(info: np.ndarray, p: ChainParams, rng) -> List[dict]:
    """Rotation step bookkeeping + draws (rotations.py:77-119).  info = stage-A rows."""
    rots = []
    for f in range(info.shape[0]):
        x, y, w, h = (int(v) for v in info[f, :4])
        if w <= 0 or h <= 0:
            continue              # the key removed everything: keep_largest_component produced no file
        if p.include_original:
            rots.append(dict(frame=f, cc=(x, y, w, h), rot_index=0, angle=None,
                             plan=hostmath.RotationPlan("copy", (w, h), None, None, True), original=True))
        for i in range(p.num_rotations):
            angle = rng.uniform(p.angle_min, p.angle_max)
            rots.append(dict(frame=f, cc=(x, y, w, h), rot_index=i + 1, angle=angle,
                             plan=hostmath.rotation_plan(w, h, angle), original=False))
    return rots


def _rot_records(rots: List[dict]):
    arr = np.zeros(len(rots), dtype=ROT_DTYPE)
    tile = 0
    for i, r in enumerate(rots):
        plan = r["plan"]
        if plan.kind == "affine" and not plan.fixed_ok:
            raise NotImplementedError("rotation outside libImaging's 16.16 fixed-point range")
        a = plan.fixed if plan.kind == "affine" else (0,) * 6
        cw, ch = plan.size
        arr[i] = (r["frame"], *r["cc"], KIND[plan.kind], *a, cw, ch, tile, 0)
        tile += ((cw + 31) // 32) * ((ch + 31) // 32)
    return arr, tile


def plan_items(rots: List[dict], bbox: np.ndarray, params: ChainParams, rng, m: int, BW: int, BH: int) -> List[dict]:
    """Host bookkeeping + draws of the steps after the rotation: getbbox fit
    (rotations.py:99-110), symmetry selection (symmetry.py:97-126), overlay size / position /
    label (overlays.py:104-149), square crop offset + label (crop_square.py:163-217).
    Pure host code (no device access): ``bbox`` holds stage B1's result per rotated sprite.
    Items are processed step-major in file-name order, drawing from ``rng`` exactly as the
    reference's steps would."""
    sprites = []
    for r, bb in zip(rots, bbox):
        cw, ch = r["plan"].size
        if r["original"] or bb[2] <= 0:          # original is saved un-cropped; empty bbox => un-cropped canvas
            fit = (0, 0, cw, ch)
        else:
            fit = (int(bb[0]), int(bb[1]), int(bb[2] - bb[0]), int(bb[3] - bb[1]))
        sprites.append(dict(r, fit=fit))

    pool = params.pool if params.pool else ["o", "h", "v", "hv"]
    if any(s not in SYM_ORDER for s in pool):
        raise ValueError(f"`pool` contient des éléments invalides : {[s for s in pool if s not in SYM_ORDER]}")
    choose = len(pool) if params.choose_random is None else params.choose_random
    if choose < 0:
        raise ValueError("`choose_random` doit être >= 0")
    variants = []
    for s in sprites:
        filt = rng.sample(pool, choose)
        if params.sym_include_original and "o" not in set(filt):
            filt.append("o")
        for sym in sorted(set(filt), key=SYM_ORDER.get):       # next step lists the directory sorted by name
            variants.append(dict(s, sym=sym))

    planned = []
    for k, v in enumerate(variants):
        bg = k % m
        ratio = rng.uniform(params.scale_min, params.scale_max)
        sw, sh = v["fit"][2], v["fit"][3]
        try:
            nw, nh = hostmath.overlay_size(sw, sh, BW, BH, ratio)
            if nw <= 0 or nh <= 0:
                raise ValueError("height and width must be > 0")
            pos_x = rng.randint(0, BW - nw)
            pos_y = rng.randint(0, BH - nh)
        except ValueError:
            continue                                            # overlays.py:151-153: print + return None
        box = np.array([pos_x, pos_y, pos_x + nw, pos_y + nh]).reshape(1, 4)
        y = np.empty_like(box, dtype=np.float32)               # xyxy2xywhn, float32 store (overlays.py:146)
        y[..., 0] = ((box[..., 0] + box[..., 2]) / 2) / BW
        y[..., 1] = ((box[..., 1] + box[..., 3]) / 2) / BH
        y[..., 2] = (box[..., 2] - box[..., 0]) / BW
        y[..., 3] = (box[..., 3] - box[..., 1]) / BH
        cx, cy, wn, hn = y[0]
        label1 = f"{params.yolo_class_id} {cx:.6f} {cy:.6f} {wn:.6f} {hn:.6f}"
        planned.append(dict(v, bg=bg, ratio=ratio, new=(nw, nh), pos=(pos_x, pos_y), label1=label1))

    finals = []
    for it in planned:
        data = np.loadtxt([it["label1"]], ndmin=2)             # crop_square.py:62-65
        cls, boxes = data[:, 0].astype(int), data[:, 1:5].astype(float)
        try:
            crop_size, (lbx, ubx), (lby, uby) = hostmath.square_crop_bounds(boxes, BW, BH)
        except RuntimeError:
            continue
        x0 = rng.randint(lbx, ubx)
        y0 = rng.randint(lby, uby)
        absolute = hostmath.xywhn2xyxy_f32(boxes, BW, BH)
        shifted = absolute - np.array([[x0, y0, x0, y0]])
        clipped = np.clip(shifted, 0, crop_size)
        valid = (clipped[:, 0] < clipped[:, 2]) & (clipped[:, 1] < clipped[:, 3])
        if not valid.any():
            continue
        kept = clipped[valid]
        nb = np.empty_like(kept, dtype=np.float32)
        nb[..., 0] = ((kept[..., 0] + kept[..., 2]) / 2) / crop_size
        nb[..., 1] = ((kept[..., 1] + kept[..., 3]) / 2) / crop_size
        nb[..., 2] = (kept[..., 2] - kept[..., 0]) / crop_size
        nb[..., 3] = (kept[..., 3] - kept[..., 1]) / crop_size
        finals.append(dict(it, sq=(x0, y0, crop_size), label=hostmath.format_label_rows(cls[valid], nb)))

    return finals


def run_chain(frames: torch.Tensor, backgrounds: torch.Tensor, params: ChainParams, rng=_random, _shard=None):
    """frames: uint8 [N, H, W, 3] BGR (cv2 order) on a CUDA device; backgrounds: uint8 [M, BH, BW, 3] RGB.
    ``_shard`` (set by sharding.run_chain_sharded): this call holds frames [lo, hi) of a larger
    job; stage-A / B1 metadata is gathered across ranks, the plan is made for ALL items with
    the shared ``rng`` and only the items of the local frames are executed.  Returns then
    (ChainResult, global item indices)."""
    if not (frames.is_cuda and backgrounds.is_cuda):
        raise RuntimeError("run_chain needs CUDA tensors (there is no CPU fallback)")
    if frames.dtype != torch.uint8 or backgrounds.dtype != torch.uint8 or frames.dim() != 4 or backgrounds.dim() != 4:
        raise TypeError("frames / backgrounds must be uint8 [N, H, W, 3]")
    lib = _native.load()
    dev = frames.device
    frames = frames.contiguous()
    backgrounds = backgrounds.contiguous()
    n, FH, FW, _ = frames.shape
    m, BH, BW, _ = backgrounds.shape
    rect = hostmath.border_crop_rect(FW, FH, params.crop_margins)
    ranges = hostmath.pack_hsv_ranges(params.color_ranges_to_exclude_hsv, params.zones, params.use_gimp_scale)
    geom = _geom_struct(FW, FH, rect, BW, BH, ranges)
    cs = min(BH, BW)

    with torch.cuda.device(dev):
        # ------------------------------------------------------------------ stage A
        ws_bytes = lib.ipp_chain_mask_workspace_bytes(C.byref(geom), n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        info_d = torch.empty((n, 8), dtype=torch.int32, device=dev)
        _native.check(lib.ipp_chain_stage_a(frames.data_ptr(), FW * 3, FH * FW * 3, C.byref(geom), info_d.data_ptr(),
                                            ws.data_ptr(), ws_bytes, n, _stream()), "ipp_chain_stage_a")
        info = info_d.cpu().numpy()                                   # sync #1
        lo = 0
        if _shard is not None:
            lo = _shard["lo"]
            info = _shard["gather"](info)                             # 32 bytes of metadata per frame

        # ------------------------------------------------------------------ rotation step (host) + stage B1
        rots = plan_rotations(info, params, rng)
        if not rots and _shard is None:
            return ChainResult(torch.empty((0, cs, cs, 3), dtype=torch.uint8, device=dev), [], [])
        all_rots = rots
        if _shard is not None:                                        # execute only the sprites of local frames
            rots = [dict(r, frame=r["frame"] - lo) for r in all_rots if lo <= r["frame"] < _shard["hi"]]
        rot_arr, rot_tiles = _rot_records(rots)
        rot_d = torch.from_numpy(rot_arr.view(np.uint8)).to(dev)
        bbox_d = torch.empty((len(rots), 4), dtype=torch.int32, device=dev)
        if len(rots):
            _native.check(lib.ipp_chain_stage_b1(C.byref(geom), rot_d.data_ptr(), len(rots), rot_tiles, ws.data_ptr(),
                                                 bbox_d.data_ptr(), _stream()), "ipp_chain_stage_b1")
        bbox = bbox_d.cpu().numpy()                                   # sync #2
        if _shard is not None:
            bbox = _shard["gather"](bbox)                             # 16 bytes per sprite, rank order == item order
            rots = all_rots

        finals = plan_items(rots, bbox, params, rng, m, BW, BH)

        own = list(range(len(finals)))
        if _shard is not None:
            own = [i for i, f in enumerate(finals) if lo <= f["frame"] < _shard["hi"]]
            finals = [dict(finals[i], frame=finals[i]["frame"] - lo) for i in own]
        K = len(finals)
        out = torch.empty((K, cs, cs, 3), dtype=torch.uint8, device=dev)
        if K == 0:
            return ChainResult(out, [], []) if _shard is None else (ChainResult(out, [], []), own)

        # ------------------------------------------------------------------ item records, tables, sprite workspace
        items = np.zeros(K, dtype=ITEM_DTYPE)
        tables: List[np.ndarray] = []
        toff = 0
        woff = 0
        gt = ht = vt = 0

        def align(v, a=256):
            return (v + a - 1) // a * a

        for i, it in enumerate(finals):
            plan = it["plan"]
            a = plan.fixed if plan.kind == "affine" else (0,) * 6
            fx, fy, sw, sh = it["fit"]
            nw, nh = it["new"]
            rec = items[i]
            rec["frame"], rec["bg"] = it["frame"], it["bg"]
            rec["cc_x"], rec["cc_y"], rec["cc_w"], rec["cc_h"] = it["cc"]
            rec["kind"] = KIND[plan.kind]
            for j in range(6):
                rec[f"a{j}"] = a[j]
            rec["fit_x0"], rec["fit_y0"], rec["spr_w"], rec["spr_h"] = fx, fy, sw, sh
            rec["flip_h"] = int(it["sym"] in ("h", "hv"))
            rec["flip_v"] = int(it["sym"] in ("v", "hv"))
            rec["new_w"], rec["new_h"] = nw, nh
            rec["pos_x"], rec["pos_y"] = it["pos"]
            rec["sq_x0"], rec["sq_y0"], rec["sq_size"] = it["sq"]
            if nw != sw:
                ks = lib.ipp_lanczos_ksize(sw, nw)
                b = np.empty((nw, 2), dtype=np.int32)
                kk = np.empty((nw, ks), dtype=np.int32)
                _native.check(lib.ipp_lanczos_coeffs_host(sw, nw, b.ctypes.data, kk.ctypes.data), "ipp_lanczos_coeffs_host")
                rec["h_ksize"], rec["hb_off"], rec["hk_off"] = ks, toff, toff + b.size
                tables += [b.ravel(), kk.ravel()]
                toff += b.size + kk.size
            if nh != sh:
                ks = lib.ipp_lanczos_ksize(sh, nh)
                b = np.empty((nh, 2), dtype=np.int32)
                kk = np.empty((nh, ks), dtype=np.int32)
                _native.check(lib.ipp_lanczos_coeffs_host(sh, nh, b.ctypes.data, kk.ctypes.data), "ipp_lanczos_coeffs_host")
                rec["v_ksize"], rec["vb_off"], rec["vk_off"] = ks, toff, toff + b.size
                tables += [b.ravel(), kk.ravel()]
                toff += b.size + kk.size
            rec["gather_tile_start"], rec["h_tile_start"], rec["v_tile_start"] = gt, ht, vt
            gt += ((sw + 31) // 32) * ((sh + 31) // 32)
            ht += ((nw + 31) // 32) * ((sh + 7) // 8)
            vt += ((nw + 31) // 32) * ((nh + 7) // 8)
            rec["rot_off"] = woff; woff = align(woff + sw * sh * 4)
            rec["tmp_off"] = woff; woff = align(woff + nw * sh * 4)
            rec["spr_off"] = woff; woff = align(woff + nw * nh * 4)

        items_d = torch.from_numpy(items.view(np.uint8)).to(dev)
        tab = np.concatenate(tables) if tables else np.zeros(1, dtype=np.int32)
        tab_d = torch.from_numpy(tab).to(dev)
        spr_ws = torch.empty(max(woff, 256), dtype=torch.uint8, device=dev)

        # ------------------------------------------------------------------ stage B2 + C
        _native.check(lib.ipp_chain_stage_b2(frames.data_ptr(), FW * 3, FH * FW * 3, C.byref(geom), items_d.data_ptr(), K,
                                             gt, ht, vt, tab_d.data_ptr(), ws.data_ptr(), spr_ws.data_ptr(), _stream()),
                      "ipp_chain_stage_b2")
        _native.check(lib.ipp_chain_stage_c(backgrounds.data_ptr(), BW * 3, BH * BW * 3, C.byref(geom), items_d.data_ptr(), K,
                                            spr_ws.data_ptr(), out.data_ptr(), cs * 3, cs * cs * 3, cs, _stream()),
                      "ipp_chain_stage_c")
        # keep the scratch alive until the stream has consumed it
        for t in (ws, rot_d, items_d, tab_d, spr_ws):
            t.record_stream(torch.cuda.current_stream())

    records = [dict(frame=it["frame"], rot_index=it["rot_index"], angle=it["angle"], sym=it["sym"], bg=it["bg"],
                    ratio=it["ratio"], new=it["new"], pos=it["pos"], sq=it["sq"], fit=it["fit"], cc=it["cc"]) for it in finals]
    res = ChainResult(out, [it["label"] for it in finals], records)
    return res if _shard is None else (res, own)


def run_chain_host(frames_host: torch.Tensor, backgrounds_host: torch.Tensor, params: ChainParams, device,
                   out_host: Optional[torch.Tensor] = None, rng=_random) -> ChainResult:
    """End-to-end entry for HOST buffers (the call a pipeline step makes after decoding):
    pinned host frames / backgrounds -> H2D -> run_chain -> D2H of the output images.
    ``out_host`` (pinned, [>=K, cs, cs, 3]) receives the images; ChainResult.images is that
    host tensor (first K rows)."""
    dev = torch.device(device)
    with torch.cuda.device(dev):
        f = frames_host.to(dev, non_blocking=True)
        b = backgrounds_host.to(dev, non_blocking=True)
        res = run_chain(f, b, params, rng)
        k = res.images.shape[0]
        if out_host is None:
            out_host = torch.empty(res.images.shape, dtype=torch.uint8).pin_memory()
        out_host[:k].copy_(res.images, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return ChainResult(out_host[:k], res.labels, res.records)
