"""Oracle of the full chain: the reference's steps, one after the other, with the
reference's own library calls (oracle.libcalls, T2) on host arrays.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Mirrors what
``ProcessingPipeline.run()`` (reference pipeline.py:555-566) does when the steps
crop_from_border -> process_images_with_color_masks -> keep_largest_component ->
process_rotations -> generate_symmetries -> paste_overlay_onto_background ->
process_square_crop_around_bbox are chained: every step runs over all items in file-name
order, draws from the module-level ``random``, and items that fail disappear.  Files are
replaced by in-memory arrays (PNG is lossless; the JPEG hops of the real workflow are
codecs, outside the hot path).
"""
from __future__ import annotations

import random

import cv2
import numpy as np

from . import libcalls as L

SYM_FILE_ORDER = {"h": 0, "hv": 1, "o": 2, "v": 3}


def run_chain(frames_bgr, backgrounds_rgb, *, crop_margins=(0, 0, 0, 0), color_ranges_to_exclude_hsv=(), zones=None,
              use_gimp_scale=False, num_rotations=10, include_original=True, angle_min=1.0, angle_max=359.0,
              pool=None, choose_random=None, sym_include_original=True, yolo_class_id=0, scale_min=0.15,
              scale_max=0.30, fast_cc=True, enhance=False):
    """Returns (images [list of HxWx3 RGB arrays], labels [list of str], records [list of dict])."""
    # step 1+2+3: border crop, HSV key, largest component (+ fit)
    sprites = []
    for f, frame in enumerate(frames_bgr):
        cropped = L.crop_from_border(frame, crop_margins)
        keyed = L.color_masks_to_alpha(np.ascontiguousarray(cropped), list(color_ranges_to_exclude_hsv), zones, use_gimp_scale)
        try:
            fitted = L.keep_largest_component(keyed)
        except cv2.error:
            continue
        if fitted.size == 0:
            continue                      # cv2.imwrite of an empty image raises -> caught -> None (pixels_isolés.py:65-72)
        sprites.append(dict(frame=f, img=np.ascontiguousarray(fitted[..., [2, 1, 0, 3]])))   # BGRA file == RGBA for PIL

    # step 4: rotations (rotations.py:77-119)
    rotated = []
    for s in sprites:
        rgba = L.to_rgba(s["img"])
        if include_original:
            rotated.append(dict(frame=s["frame"], rot_index=0, angle=None, img=rgba))
        for i in range(num_rotations):
            angle = random.uniform(angle_min, angle_max)
            img, _ = L.rotate_and_fit(rgba, angle)
            rotated.append(dict(frame=s["frame"], rot_index=i + 1, angle=angle, img=img))

    # step 5: symmetries (symmetry.py:97-137); the next step sees the files sorted by name
    variants = []
    for r in rotated:
        syms = L.select_symmetries(pool, choose_random, sym_include_original)
        imgs = dict(zip(syms, L.symmetries(r["img"], syms)))
        for sym in sorted(imgs, key=SYM_FILE_ORDER.get):
            variants.append(dict(r, sym=sym, img=imgs[sym]))

    # step 6: overlay onto background, 'modulo' pairing (overlays.py, pipeline.py:204-207)
    composites = []
    m = len(backgrounds_rgb)
    for k, v in enumerate(variants):
        bg = backgrounds_rgb[k % m]
        try:
            comp, label, (nw, nh, px, py) = L.overlay_onto_background(v["img"], bg, yolo_class_id, scale_min, scale_max)
        except ValueError:
            continue
        composites.append(dict(v, bg=k % m, img=comp, label=label, new=(nw, nh), pos=(px, py)))

    # step 7: square crop around the bbox (crop_square.py, repaired)
    images, labels, records = [], [], []
    for c in composites:
        data = np.loadtxt([c["label"]], ndmin=2)
        cls, boxes = data[:, 0].astype(int), data[:, 1:5].astype(float)
        try:
            cropped, ncls, nboxes, (x0, y0) = L.square_crop(c["img"], cls, boxes)
        except RuntimeError:
            continue
        images.append(cropped)
        labels.append(L.format_label_rows(ncls, nboxes))
        records.append(dict(frame=c["frame"], rot_index=c["rot_index"], angle=c["angle"], sym=c["sym"], bg=c["bg"],
                            new=c["new"], pos=c["pos"], sq=(x0, y0, cropped.shape[0])))

    # step 8 (optional): photometric jitter, 'sample' pairing (pipeline.py:209-225, tranfo.py:37-51)
    if enhance and images:
        do_blur, do_rgb = L.sample_pairing_flags(len(images))
        images = [L.enhance_image(np.ascontiguousarray(img), bl, rg) for img, bl, rg in zip(images, do_blur, do_rgb)]
    return images, labels, records
