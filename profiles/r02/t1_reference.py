"""T1 baseline lines: the REAL reference functions (/root/reference, imported by oracle.ref_loader) driven by
the reference's own ProcessingStep over PNG / JPG files on tmpfs, all host cores (workers=-1, pipeline.py:88-90,372).
Shows where the reference's wall time goes (codecs) next to the in-memory oracle numbers of bench.py --impl reference.
Build container only (the GPU box has no /root/reference):   python profiles/r02/t1_reference.py > /dev/null     (writes profiles/r02/t1_reference_lines.jsonl)
"""
import json, os, random, shutil, sys, tempfile, time
from pathlib import Path

import cv2
import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from image_processor_pipeline_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = open(Path(__file__).resolve().parent / "t1_reference_lines.jsonl", "w")
ref = ref_loader.load()
P = ref.pipeline
base = Path(tempfile.mkdtemp(prefix="ipp_t1_", dir="/dev/shm"))
cores = os.cpu_count()


def run_steps(name, steps, n_images, prepare_s):
    random.seed(1)
    t0 = time.perf_counter()
    for st in steps:
        st.run()
    dt = time.perf_counter() - t0
    outs = sum(1 for p in Path(steps[-1].output_paths[0]).iterdir())
    OUT.write(json.dumps({"impl": "reference-T1", "config": name, "value": n_images / dt, "unit": "input images/s", "seconds": dt,
                      "inputs": n_images, "files_written_by_last_step": outs, "cores": cores, "where": "build container (no GPU), tmpfs",
                      "what": "the reference's own functions through ProcessingStep(workers=-1), decode + arithmetic + encode of PNG/JPG files",
                      "input_generation_s": prepare_s}) + "\n")
    OUT.flush()


def step(name, fn, ins, outs, workers=-1, **kw):
    return P.ProcessingStep(name, fn, input_dirs=ins, output_dirs=outs, root_dir=base, workers=workers, **kw)


def fresh(*names):
    out = []
    for n in names:
        d = base / n
        if d.exists():
            shutil.rmtree(d)
        d.mkdir(parents=True)
        out.append(d)
    return out


# ---- c1: symmetry + crop_square (1024 x 1024 RGB PNG + one YOLO box per image)
N = 32
t0 = time.perf_counter()
d_in, d_lab = fresh("c1/in", "c1/lab")
for i, img in enumerate(bench.noise_images_host(N, bench.C1_SHAPE, 1, 0)):
    cv2.imwrite(str(d_in / f"img_{i:04d}.png"), img)
prep = time.perf_counter() - t0
s1 = step("sym", ref.symmetry.generate_symmetries, ["c1/in"], ["c1/sym"])
run_steps("c1 symmetry (crop_square needs labels per variant: not chained by the reference either)", [s1], N, prep)

# ---- c2: rotations (1080p RGB PNG, one rotation per image)
N = 16
t0 = time.perf_counter()
(d_in,) = fresh("c2/in")
for i, img in enumerate(bench.noise_images_host(N, bench.C2_SHAPE, 1, 0)):
    Image.fromarray(img, "RGB").save(d_in / f"img_{i:04d}.png")
prep = time.perf_counter() - t0
run_steps("c2 rotations", [step("rot", ref.rotations.process_rotations, ["c2/in"], ["c2/out"], options=dict(num_rotations=1, include_original=False))], N, prep)

# ---- c3: overlays (512 x 384 RGBA sprites onto 1080p backgrounds, 'modulo' pairing)
N = 32
t0 = time.perf_counter()
d_spr, d_bg = fresh("c3/spr", "c3/bg")
for i, s in enumerate(bench.sprite_images(N, bench.SPR[0], bench.SPR[1], 1)):
    Image.fromarray(s, "RGBA").save(d_spr / f"spr_{i:04d}.png")
for i, img in enumerate(bench.noise_images_host(8, bench.C2_SHAPE, 1, 0)):
    Image.fromarray(img, "RGB").save(d_bg / f"bg_{i:04d}.png")
prep = time.perf_counter() - t0
run_steps("c3 overlays", [step("ov", ref.overlays.paste_overlay_onto_background, ["c3/spr", "c3/bg"], ["c3/img", "c3/lab"], pairing_method="modulo")], N, prep)

# ---- c4: filtres_liste + pixels_isolés (512 x 512 JPG frames -> BGRA PNG)
N = 128
t0 = time.perf_counter()
(d_in,) = fresh("c4/in")
for i in range(N):
    cv2.imwrite(str(d_in / f"f_{i:04d}.png"), synth.frame(1, i % 16, 512, 512))
prep = time.perf_counter() - t0
s1 = step("key", ref.filtres_liste.process_images_with_color_masks, ["c4/in"], ["c4/keyed"], options=dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS))
s2 = step("isoles", ref.pixels_isoles.keep_largest_component, ["c4/keyed"], ["c4/out"])
run_steps("c4 filtres_liste -> pixels_isolés", [s1, s2], N, prep)

# ---- headline chain on 1080p frames: the seven steps through files (frames as JPG like video.py writes them)
N = 8
t0 = time.perf_counter()
d_fr, d_bg = fresh("chain/0-raw", "chain/bg")
frames, bgs = bench.make_inputs(N, 1, 0, base=4)
for i in range(N):
    cv2.imwrite(str(d_fr / f"obj-frame_{i:04d}.jpg"), frames[i])
    Image.fromarray(bgs[i], "RGB").save(d_bg / f"bg_{i:04d}.png")
prep = time.perf_counter() - t0
kw = bench.CHAIN_KW
steps = [
    step("crop", ref.recadrages.crop_from_border, ["chain/0-raw"], ["chain/1-crop"]),
    step("key", ref.filtres_liste.process_images_with_color_masks, ["chain/1-crop"], ["chain/2-key"], options=dict(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS)),
    step("isoles", ref.pixels_isoles.keep_largest_component, ["chain/2-key"], ["chain/3-cc"]),
    step("rot", ref.rotations.process_rotations, ["chain/3-cc"], ["chain/4-rot"], options=dict(num_rotations=kw["num_rotations"], include_original=kw["include_original"])),
    step("sym", ref.symmetry.generate_symmetries, ["chain/4-rot"], ["chain/5-sym"], options=dict(pool=kw["pool"], choose_random=kw["choose_random"], include_original=kw["sym_include_original"])),
    step("ov", ref.overlays.paste_overlay_onto_background, ["chain/5-sym", "chain/bg"], ["chain/6-img", "chain/6-lab"], pairing_method="modulo"),
    step("sq", ref.crop_square.process_square_crop_around_bbox, ["chain/6-img", "chain/6-lab"], ["chain/7-img", "chain/7-lab"], pairing_method="zip",
         workers=1),     # the repaired copy of crop_square.py is loaded from a temp file and cannot be pickled for the process pool
]
run_steps("chain (7 steps through files; JPG frames are lossy, so pixels differ from the in-memory oracle: timing only)", steps, N, prep)
shutil.rmtree(base, ignore_errors=True)
