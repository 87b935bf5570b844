"""Generate the committed golden fixtures by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every transform on the hot path it writes small deterministic inputs to a
temp dir, seeds ``random``, calls the reference's own function (imported by
``oracle.ref_loader`` behind the documented shims; ``crop_square`` with the
three one-token repairs of SURVEY.md §8a-4), reads the written files back and
stores inputs, seeds, options, output names and decoded output arrays in
``tests/golden/<case>.npz``.  The fixtures pin (a) the oracle (CPU tests) and
(b) the CUDA drop-in functions (GPU tests).  Library versions are recorded in
``tests/golden/VERSIONS.json``.
"""
from __future__ import annotations

import io
import json
import random
import sys
import tempfile
from pathlib import Path

import cv2
import numpy as np
import PIL
from PIL import Image

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

OUT = Path(__file__).resolve().parent


def _rng(seed):
    return np.random.default_rng(seed)


def blob_alpha(h, w, seed, specks=0.03):
    """Alpha plane with a few blobs + specks (values 0, 1 and >1 present)."""
    rng = _rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.zeros((h, w), np.uint8)
    for _ in range(3):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        ry, rx = rng.integers(3, max(4, h // 3)), rng.integers(3, max(4, w // 3))
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        a[m] = rng.integers(2, 256)
    sp = rng.random((h, w)) < specks
    a[sp] = rng.integers(0, 256, (h, w), dtype=np.uint8)[sp]
    a[rng.random((h, w)) < 0.01] = 1
    return a


def read_unchanged(p):
    return cv2.imread(str(p), cv2.IMREAD_UNCHANGED)


def save(case, **kw):
    meta = kw.pop("meta")
    np.savez_compressed(OUT / f"{case}.npz", meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **kw)
    print(f"  {case}.npz  ({(OUT / (case + '.npz')).stat().st_size} B)")


def main():
    ref = ref_loader.load()
    tmp = Path(tempfile.mkdtemp(prefix="ipp_golden_"))
    versions = {"cv2": cv2.__version__, "PIL": PIL.__version__, "numpy": np.__version__,
                "python": sys.version.split()[0], "crop_square_repairs": [list(r) for r in ref_loader.CROP_SQUARE_REPAIRS]}
    (OUT / "VERSIONS.json").write_text(json.dumps(versions, indent=1))

    # ---------------------------------------------------------------- symmetry
    for k, (shape, seed, opts) in enumerate([
        ((37, 53, 3), 11, {}),
        ((20, 31, 4), 12, {"pool": ["h", "v", "hv"], "choose_random": 2, "include_original": True}),
        ((16, 16), 13, {"pool": ["hv", "o"], "include_original": False}),
    ]):
        d = tmp / f"sym{k}"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
        img = _rng(seed).integers(0, 256, shape, dtype=np.uint8)
        cv2.imwrite(str(d / "in" / "img.png"), img)
        random.seed(seed)
        res = ref.symmetry.generate_symmetries(d / "in" / "img.png", [d / "out"], **opts)
        outs = {f"out_{i}": read_unchanged(p) for i, p in enumerate(res)}
        save(f"symmetry_{k}", input=img, meta={"seed": seed, "opts": opts, "names": [p.name for p in res]}, **outs)

    # ---------------------------------------------------------------- rotations
    for k, (shape, seed, opts) in enumerate([
        ((31, 40, 3), 21, {"num_rotations": 3}),
        ((24, 18, 4), 22, {"num_rotations": 2, "include_original": False, "angle_min": 80.0, "angle_max": 100.0}),
    ]):
        d = tmp / f"rot{k}"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
        img = _rng(seed).integers(0, 256, shape, dtype=np.uint8)
        if shape[2] == 4:
            img[..., 3] = blob_alpha(shape[0], shape[1], seed)
        Image.fromarray(img, "RGB" if shape[2] == 3 else "RGBA").save(d / "in" / "spr.png")
        random.seed(seed)
        res = ref.rotations.process_rotations(d / "in" / "spr.png", [d / "out"], **opts)
        outs = {f"out_{i}": np.asarray(Image.open(p)) for i, p in enumerate(res)}
        random.seed(seed)
        angles = [random.uniform(opts.get("angle_min", 1.0), opts.get("angle_max", 359.0)) for _ in range(opts["num_rotations"])]
        save(f"rotations_{k}", input=img, angles=np.array(angles),
             meta={"seed": seed, "opts": opts, "names": [p.name for p in res]}, **outs)

    # ---------------------------------------------------------------- tranfo (photometric jitter, pairing mode 'sample')
    for k, (shape, seed, blur, rgb) in enumerate([((48, 70, 3), 61, False, False), ((40, 33, 3), 62, True, False),
                                                  ((29, 64, 3), 63, False, True), ((56, 56, 3), 64, True, True)]):
        d = tmp / f"tranfo{k}"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
        img = _rng(seed).integers(0, 256, shape, dtype=np.uint8)
        Image.fromarray(img, "RGB").save(d / "in" / "img.png")
        random.seed(seed)
        res = ref.tranfo.enhance_image(d / "in" / "img.png", blur, rgb, [d / "out"])
        save(f"tranfo_{k}", input=img, out=np.asarray(Image.open(res).convert("RGB")),
             meta={"seed": seed, "apply_blur": blur, "apply_rgb": rgb, "name": res.name})

    # ---------------------------------------------------------------- crop_square (repaired copy)
    for k, (shape, seed, labels) in enumerate([
        ((50, 80, 3), 31, "0 0.5 0.5 0.2 0.3\n1 0.55 0.4 0.1 0.1\n"),
        ((90, 60, 3), 32, "3 0.4 0.6 0.25 0.2\n"),
        ((64, 64, 3), 33, "2 0.3 0.3 0.1 0.1\n7 0.8 0.7 0.2 0.2\n"),
    ]):
        d = tmp / f"sq{k}"; (d / "in").mkdir(parents=True); (d / "o1").mkdir(); (d / "o2").mkdir()
        img = _rng(seed).integers(0, 256, shape, dtype=np.uint8)
        cv2.imwrite(str(d / "in" / "a.png"), img)
        (d / "in" / "a.txt").write_text(labels)
        random.seed(seed)
        res = ref.crop_square.process_square_crop_around_bbox(d / "in" / "a.png", d / "in" / "a.txt", [d / "o1", d / "o2"])
        save(f"crop_square_{k}", input=img, out_image=read_unchanged(res[0]),
             meta={"seed": seed, "labels_in": labels, "labels_out": res[1].read_text(), "names": [p.name for p in res]})

    # ---------------------------------------------------------------- crop_from_border / fit_crop
    d = tmp / "border"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
    img = cv2.GaussianBlur(_rng(41).integers(0, 256, (60, 90, 3), dtype=np.uint8), (7, 7), 0)
    cv2.imwrite(str(d / "in" / "f.jpg"), img)
    jpg_bytes = np.frombuffer((d / "in" / "f.jpg").read_bytes(), dtype=np.uint8)
    margins = (0.1, 5, 0.25, 7)
    res = ref.recadrages.crop_from_border(d / "in" / "f.jpg", [d / "out"], crop_margins=margins)
    save("crop_from_border_0", input_jpg=jpg_bytes, input_decoded=read_unchanged(d / "in" / "f.jpg"),
         out_decoded=read_unchanged(res), out_jpg=np.frombuffer(res.read_bytes(), dtype=np.uint8),
         meta={"margins": margins, "name": res.name})

    d = tmp / "fit"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
    img = _rng(42).integers(0, 256, (40, 50, 4), dtype=np.uint8)
    img[..., 3] = 0
    img[7:29, 11:43, 3] = _rng(43).integers(0, 256, (22, 32), dtype=np.uint8)
    img[7, 11:43, 3] = 0; img[8, 20, 3] = 9
    Image.fromarray(img, "RGBA").save(d / "in" / "s.png")
    res = ref.recadrages.fit_crop(d / "in" / "s.png", [d / "out"])
    save("fit_crop_0", input=img, out=np.asarray(Image.open(res)), meta={"name": res.name})

    # ---------------------------------------------------------------- overlays
    for k, (sshape, bshape, seed, opts) in enumerate([
        ((45, 60, 4), (120, 160, 3), 51, {}),
        ((70, 30, 4), (100, 90, 3), 52, {"yolo_class_id": 4, "scale_min": 0.4, "scale_max": 0.9}),
        ((20, 25, 4), (200, 150, 3), 53, {"scale_min": 0.5, "scale_max": 0.6}),   # up-scaling
    ]):
        d = tmp / f"ov{k}"; (d / "spr").mkdir(parents=True); (d / "bg").mkdir(); (d / "o1").mkdir(); (d / "o2").mkdir()
        spr = _rng(seed).integers(0, 256, sshape, dtype=np.uint8)
        spr[..., 3] = blob_alpha(sshape[0], sshape[1], seed, specks=0.1)
        bg = _rng(seed + 100).integers(0, 256, bshape, dtype=np.uint8)
        Image.fromarray(spr, "RGBA").save(d / "spr" / "obj.png")
        Image.fromarray(bg, "RGB").save(d / "bg" / "back.png")
        random.seed(seed)
        res = ref.overlays.paste_overlay_onto_background(d / "spr" / "obj.png", d / "bg" / "back.png", [d / "o1", d / "o2"], **opts)
        save(f"overlays_{k}", sprite=spr, background=bg, out_image=np.asarray(Image.open(res[0])),
             meta={"seed": seed, "opts": opts, "label": res[1].read_text(), "names": [p.name for p in res]})

    # ---------------------------------------------------------------- pixels_isolés
    for k, (shape, seed) in enumerate([((64, 64), 61), ((45, 77), 62), ((33, 20), 63)]):
        d = tmp / f"cc{k}"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
        img = _rng(seed).integers(0, 256, shape + (4,), dtype=np.uint8)
        img[..., 3] = blob_alpha(shape[0], shape[1], seed, specks=0.08)
        cv2.imwrite(str(d / "in" / "m.png"), img)
        res = ref.pixels_isoles.keep_largest_component(d / "in" / "m.png", [d / "out"])
        save(f"pixels_isoles_{k}", input=img, out=read_unchanged(res), meta={"name": res.name})

    # ---------------------------------------------------------------- filtres_liste
    filters = [(0, 0, 0, 180, 255, 150), (15, 60, 200, 35, 255, 255),
               (15, 30 * 2.55, 55 * 2.55, 30, 60 * 2.55, 80 * 2.55),
               (15, 60 * 2.55, 60 * 2.55, 30, 75 * 2.55, 90 * 2.55)]      # reference filtres_liste.py:186-190
    gimp_filters = [(40, 20, 30, 120, 90, 100), (300, 0, 50, 360, 100, 100)]
    for k, (shape, seed, fl, zones, gimp, prefix) in enumerate([
        ((48, 64, 3), 71, filters, None, False, ""),
        ((30, 41, 3), 72, filters, [(0, 0, 0, 0), (3, 4, 5, 6), None, (10, 0, 0, 20)], False, "pre"),
        ((25, 25, 3), 73, gimp_filters, None, True, ""),
    ]):
        d = tmp / f"fl{k}"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
        img = _rng(seed).integers(0, 256, shape, dtype=np.uint8)
        cv2.imwrite(str(d / "in" / "frame.png"), img)
        res = ref.filtres_liste.process_images_with_color_masks(
            d / "in" / "frame.png", [d / "out"], color_ranges_to_exclude_hsv=fl, zones=zones,
            use_gimp_scale=gimp, output_prefix=prefix)
        save(f"filtres_liste_{k}", input=img, out=read_unchanged(res),
             meta={"filters": fl, "zones": zones, "gimp": gimp, "prefix": prefix, "name": res.name})

    # ---------------------------------------------------------------- labels (reference labels.py:77-117 known answer)
    d = tmp / "lab"; (d / "in").mkdir(parents=True); (d / "out").mkdir()
    text = "0 0.5 0.5 0.2 0.2\n1 0.3 0.3 0.1 0.1\n\n2 0.7 0.7 0.15 0.15\n0 0.1 0.1 0.05 0.05\n"
    (d / "in" / "l.txt").write_text(text)
    res = ref.labels.change_label_class(d / "in" / "l.txt", [d / "out"], cls_mapping={0: 99, 1: 77})
    save("labels_0", meta={"text_in": text, "text_out": res.read_text(), "mapping": {"0": 99, "1": 77}})

    # ---------------------------------------------------------------- RNG contract: draw order per transform
    random.seed(1234)
    draws = {"uniform_1_359": [random.uniform(1.0, 359.0) for _ in range(3)],
             "sample4": random.sample(["o", "h", "v", "hv"], 2),
             "randint_0_1000": [random.randint(0, 1000) for _ in range(3)]}
    (OUT / "rng_contract.json").write_text(json.dumps(draws, indent=1))
    print("done")


if __name__ == "__main__":
    main()
