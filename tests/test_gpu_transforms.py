"""The file-based drop-in process functions against outputs of the REAL reference
functions (tests/golden, produced by tests/golden/make_golden.py): same inputs, same
seeds, same options -> same file names, same pixels, same label text."""
import random
from pathlib import Path

import cv2
import numpy as np
import pytest
from PIL import Image

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from image_processor_pipeline_b200.pipeline import ProcessingPipeline, ProcessingStep  # noqa: E402
from image_processor_pipeline_b200.transforms import (crop_square, filtres_liste, labels, overlays, recadrages, rotations,  # noqa: E402
                                                      symmetry)
import importlib  # noqa: E402

pixels_isoles = importlib.import_module("image_processor_pipeline_b200.transforms.pixels_isolés")


def _dirs(tmp_path, *names):
    out = []
    for n in names:
        (tmp_path / n).mkdir(parents=True, exist_ok=True)
        out.append(tmp_path / n)
    return out


@pytest.mark.parametrize("k", [0, 1, 2])
def test_generate_symmetries(tmp_path, k):
    g = load_golden(f"symmetry_{k}")
    din, dout = _dirs(tmp_path, "in", "out")
    cv2.imwrite(str(din / "img.png"), g["input"])
    random.seed(g["meta"]["seed"])
    res = symmetry.generate_symmetries(din / "img.png", [dout], **g["meta"]["opts"])
    assert [p.name for p in res] == g["meta"]["names"]
    for i, p in enumerate(res):
        assert np.array_equal(cv2.imread(str(p), cv2.IMREAD_UNCHANGED), g[f"out_{i}"])


def test_generate_symmetries_errors(tmp_path):
    din, dout = _dirs(tmp_path, "in", "out")
    cv2.imwrite(str(din / "img.png"), np.zeros((4, 4, 3), np.uint8))
    with pytest.raises(ValueError):
        symmetry.generate_symmetries(din / "img.png", [])
    with pytest.raises(ValueError):
        symmetry.generate_symmetries(din / "img.xyz", [dout])
    with pytest.raises(ValueError):
        symmetry.generate_symmetries(din / "img.png", [dout], pool=["o", "diag"])
    with pytest.raises(ValueError):
        symmetry.generate_symmetries(din / "img.png", [dout], choose_random=-1)
    with pytest.raises(FileNotFoundError):
        symmetry.generate_symmetries(din / "absent.png", [dout])


@pytest.mark.parametrize("k", [0, 1])
def test_process_rotations(tmp_path, k):
    g = load_golden(f"rotations_{k}")
    din, dout = _dirs(tmp_path, "in", "out")
    img = g["input"]
    Image.fromarray(img, "RGB" if img.shape[2] == 3 else "RGBA").save(din / "spr.png")
    random.seed(g["meta"]["seed"])
    res = rotations.process_rotations(din / "spr.png", [dout], **g["meta"]["opts"])
    assert [p.name for p in res] == g["meta"]["names"]
    for i, p in enumerate(res):
        assert np.array_equal(np.asarray(Image.open(p)), g[f"out_{i}"])
    assert rotations.process_rotations(din / "absent.png", [dout]) is None
    assert rotations.process_rotations(din / "spr.png", []) is None


@pytest.mark.parametrize("k", [0, 1, 2])
def test_process_square_crop_around_bbox(tmp_path, k):
    g = load_golden(f"crop_square_{k}")
    din, o1, o2 = _dirs(tmp_path, "in", "o1", "o2")
    cv2.imwrite(str(din / "a.png"), g["input"])
    (din / "a.txt").write_text(g["meta"]["labels_in"])
    random.seed(g["meta"]["seed"])
    res = crop_square.process_square_crop_around_bbox(din / "a.png", din / "a.txt", [o1, o2])
    assert [p.name for p in res] == g["meta"]["names"]
    assert np.array_equal(cv2.imread(str(res[0]), cv2.IMREAD_UNCHANGED), g["out_image"])
    assert res[1].read_text() == g["meta"]["labels_out"]
    with pytest.raises(IndexError):
        crop_square.process_square_crop_around_bbox(din / "a.png", din / "a.txt", [o1])
    with pytest.raises(FileNotFoundError):
        crop_square.process_square_crop_around_bbox(din / "zz.png", din / "a.txt", [o1, o2])
    (din / "far.txt").write_text("0 0.02 0.5 0.03 0.1\n1 0.98 0.5 0.03 0.1\n")
    if g["input"].shape[1] > g["input"].shape[0] * 1.2:
        with pytest.raises(RuntimeError):
            with pytest.warns(UserWarning):
                crop_square.process_square_crop_around_bbox(din / "a.png", din / "far.txt", [o1, o2])


def test_crop_from_border_and_fit_crop(tmp_path):
    g = load_golden("crop_from_border_0")
    din, dout = _dirs(tmp_path, "in", "out")
    (din / "f.jpg").write_bytes(g["input_jpg"].tobytes())
    res = recadrages.crop_from_border(din / "f.jpg", [dout], crop_margins=tuple(g["meta"]["margins"]))
    assert res.name == g["meta"]["name"] and res.read_bytes() == g["out_jpg"].tobytes()
    with pytest.raises(ValueError):
        recadrages.crop_from_border(din / "f.png", [dout])
    with pytest.raises(ValueError):
        recadrages.crop_from_border(din / "f.jpg", [dout], crop_margins=(0.6, 0.6, 0, 0))
    with pytest.raises(FileNotFoundError):
        recadrages.crop_from_border(din / "absent.jpg", [dout])
    g = load_golden("fit_crop_0")
    Image.fromarray(g["input"], "RGBA").save(din / "s.png")
    res = recadrages.fit_crop(din / "s.png", [dout])
    assert res.name == g["meta"]["name"] and np.array_equal(np.asarray(Image.open(res)), g["out"])
    Image.fromarray(np.zeros((5, 6, 4), np.uint8), "RGBA").save(din / "empty.png")
    assert np.asarray(Image.open(recadrages.fit_crop(din / "empty.png", [dout]))).shape == (5, 6, 4)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_paste_overlay_onto_background(tmp_path, k):
    g = load_golden(f"overlays_{k}")
    dspr, dbg, o1, o2 = _dirs(tmp_path, "spr", "bg", "o1", "o2")
    Image.fromarray(g["sprite"], "RGBA").save(dspr / "obj.png")
    Image.fromarray(g["background"], "RGB").save(dbg / "back.png")
    random.seed(g["meta"]["seed"])
    res = overlays.paste_overlay_onto_background(dspr / "obj.png", dbg / "back.png", [o1, o2], **g["meta"]["opts"])
    assert [p.name for p in res] == g["meta"]["names"]
    assert np.array_equal(np.asarray(Image.open(res[0])), g["out_image"])
    assert res[1].read_text() == g["meta"]["label"]
    with pytest.raises(IndexError):
        overlays.paste_overlay_onto_background(dspr / "obj.png", dbg / "back.png", [o1])
    assert overlays.paste_overlay_onto_background(dspr / "absent.png", dbg / "back.png", [o1, o2]) is None


@pytest.mark.parametrize("k", [0, 1, 2])
def test_keep_largest_component(tmp_path, k):
    g = load_golden(f"pixels_isoles_{k}")
    din, dout = _dirs(tmp_path, "in", "out")
    cv2.imwrite(str(din / "m.png"), g["input"])
    res = pixels_isoles.keep_largest_component(din / "m.png", [dout])
    assert res.name == g["meta"]["name"] and np.array_equal(cv2.imread(str(res), cv2.IMREAD_UNCHANGED), g["out"])
    with pytest.raises(ValueError):
        pixels_isoles.keep_largest_component(din / "m.jpg", [dout])
    with pytest.raises(FileNotFoundError):
        pixels_isoles.keep_largest_component(din / "absent.png", [dout])
    cv2.imwrite(str(din / "rgb.png"), g["input"][..., :3])
    with pytest.raises(AttributeError):
        pixels_isoles.keep_largest_component(din / "rgb.png", [dout])
    with pytest.raises(TypeError):
        pixels_isoles.keep_largest_component(din / "m.png", [dout], unknown_option=1)      # no **options, like the reference
    empty = g["input"].copy(); empty[..., 3] = 0
    cv2.imwrite(str(din / "t.png"), empty)
    assert pixels_isoles.keep_largest_component(din / "t.png", [dout]) is None


@pytest.mark.parametrize("k", [0, 1, 2])
def test_process_images_with_color_masks(tmp_path, k):
    g = load_golden(f"filtres_liste_{k}")
    m = g["meta"]
    din, dout = _dirs(tmp_path, "in", "out")
    cv2.imwrite(str(din / "frame.png"), g["input"])
    zones = [tuple(z) if z else None for z in m["zones"]] if m["zones"] else None
    res = filtres_liste.process_images_with_color_masks(din / "frame.png", [dout], [tuple(f) for f in m["filters"]], zones=zones,
                                                        use_gimp_scale=m["gimp"], output_prefix=m["prefix"])
    assert res.name == m["name"] and np.array_equal(cv2.imread(str(res), cv2.IMREAD_UNCHANGED), g["out"])
    with pytest.raises(ValueError):
        filtres_liste.process_images_with_color_masks(din / "frame.png", [dout], [])
    with pytest.raises(ValueError):
        filtres_liste.process_images_with_color_masks(din / "frame.png", [dout], [(0, 0, 0, 1, 1, 1)], zones=[None, None])
    with pytest.raises(IOError):
        filtres_liste.process_images_with_color_masks(din / "absent.png", [dout], [(0, 0, 0, 1, 1, 1)])


def test_pipeline_end_to_end_on_gpu(tmp_path):
    """symmetry -> rotations through ProcessingPipeline (the reference's own two-step smoke),
    sequential and with workers=2 (thread pool for GPU-backed functions)."""
    rng = np.random.default_rng(0)
    (tmp_path / "raw").mkdir()
    for i in range(3):
        cv2.imwrite(str(tmp_path / "raw" / f"im{i}.png"), rng.integers(0, 256, (24, 30, 3), dtype=np.uint8))
    for workers in (1, 2):
        pipe = ProcessingPipeline(root_dir=tmp_path)
        pipe.add_step(ProcessingStep("sym", symmetry.generate_symmetries, "raw", f"s{workers}", workers=workers))
        pipe.add_step(ProcessingStep("rot", rotations.process_rotations, output_dirs=f"r{workers}", workers=workers,
                                     options={"num_rotations": 2}))
        random.seed(1)
        pipe.run()
        assert len(list((tmp_path / f"s{workers}").iterdir())) == 12
        assert len(list((tmp_path / f"r{workers}").iterdir())) == 36
        assert all(l["status"] == "Success" for s in pipe.steps for l in s.process_logs)
    a = cv2.imread(str(tmp_path / "s1" / "im0_h.png"), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(a, cv2.flip(cv2.imread(str(tmp_path / "raw" / "im0.png"), cv2.IMREAD_UNCHANGED), 1))


@pytest.mark.parametrize("mode", ["P", "1", "CMYK", "I;16", "F"])
def test_fit_crop_other_pil_modes(tmp_path, mode):
    """recadrages.py:71-77 takes whatever Image.open returns: getbbox + crop for every PIL mode (device bbox on the pixel
    bytes), compared with Pillow's own getbbox / crop."""
    from PIL import Image
    r = np.random.default_rng(3)
    h, w = 23, 31
    base = np.zeros((h, w), np.uint8)
    base[5:17, 7:22] = r.integers(1, 255, (12, 15))
    if mode == "P":
        im = Image.fromarray(base, "L").convert("P")
    elif mode == "1":
        im = Image.fromarray((base > 0).astype(np.uint8) * 255, "L").convert("1")
    elif mode == "CMYK":
        im = Image.merge("CMYK", [Image.fromarray(base, "L")] * 4)
    elif mode == "I;16":
        im = Image.fromarray(base.astype(np.uint16) * 200, "I;16")
    elif mode == "PA":
        im = Image.merge("LA", [Image.fromarray(np.full((h, w), 9, np.uint8), "L"), Image.fromarray(base, "L")]).convert("PA")
    else:
        im = Image.fromarray(base.astype(np.float32) * 0.5, "F")
    ext = {"F": "tiff", "CMYK": "tiff", "I;16": "png", "PA": "png"}.get(mode, "png")
    din, dout = tmp_path / "in", tmp_path / "out"
    din.mkdir(); dout.mkdir()
    im.save(din / f"x.{ext}")
    opened = Image.open(din / f"x.{ext}")
    ref = opened.crop(opened.getbbox()) if opened.getbbox() else opened.copy()
    res = recadrages.fit_crop(din / f"x.{ext}", [dout])
    out = Image.open(res)
    assert out.size == ref.size and out.mode == ref.mode
    assert np.array_equal(np.asarray(out), np.asarray(ref))
