#!/usr/bin/env python
"""bench.py — benchmark of the B200-native transform chain and of its operators at the BASELINE.json configs.

Metric (BASELINE.json): images/sec, 1080p RGB full chain, at 1/2/4/8 B200; % of HBM roofline.
A "step" = one pass of the hot path over one batch of synthetic inputs per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]
                    [--config chain|c1|c2|c3|c4|c5]

--config (default ``chain`` = the headline; the others are BASELINE.json configs[0..4]):
  chain  full chain (border crop -> HSV key -> largest component + fit -> rotate + fit -> flip -> LANCZOS
         overlay -> square crop, YOLO labels), 1920x1080 frame + background -> 1080x1080
  c5     the same chain over a 3840x2160 frame stream (backgrounds 3840x2160 -> 2160x2160)
  c1     symmetry (o, h, v, hv) + crop_square on 1024x1024 RGB images
  c2     rotations (one per image, expand + getbbox) on 1920x1080 RGB images: nearest and bilinear
  c3     overlays: LANCZOS resize of RGBA sprites + alpha paste onto 1920x1080 backgrounds, labels
  c4     filtres_liste HSV key + pixels_isolés keep-largest-component on 512x512 images

Every line carries:
* value      images/s, inputs resident in HBM, CUDA-event timed, max over ranks, all N GPUs.
* e2e        the same metric through the public host API with pinned HOST buffers: H2D + work + D2H of the
             results inside the timed region.
* roofline   dominant streaming kernel: algorithmic bytes per launch / its mean CUDA-event duration
             (events on the launching stream around every launch, ipp_timing_* hooks).
* cpu_baseline  the oracle (the reference's own cv2 / Pillow call sequence, in memory) on this box's host cores,
             bounded sample.  Rank 0, N = 1 only.
--impl reference times that CPU path with all host cores (ProcessPoolExecutor, the reference's workers=-1 mode,
pipeline.py:88-90,372) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FH, FW = 1080, 1920
CHAIN_KW = dict(num_rotations=1, include_original=False, pool=["o", "h", "v", "hv"], choose_random=1,
                sym_include_original=False)
GENERATOR = ("frames = synth.frame (dark backdrop, every byte < 120, keyed out by filter 0 of filtres_liste.py:186-190; one bright "
             "textured ellipse ~15 % of the pixels with yellow patches the other filters remove; 1 % bright salt specks), "
             "8 base frames multiplied by cyclic shifts; backgrounds = uniform noise")

ENHANCE_NOTE = "; + enhance_image (tranfo.py: brightness / contrast / colour jitter, GaussianBlur on 30 %, channel tables on 30 %) as the last step"
CONFIGS = {
    "chain": dict(metric="images/sec 1080p RGB full chain", fh=1080, fw=1920, batch=1024, e2e_frames=512,
                  workload="full chain (border-crop + HSV key + largest component + rotate + flip + LANCZOS overlay + square crop "
                           "+ labels), 1920x1080 RGB frame + 1920x1080 background -> 1080x1080; " + GENERATOR),
    "c5": dict(metric="images/sec 4K RGB full chain (video.py frame stream)", fh=2160, fw=3840, batch=128, e2e_frames=64,
               workload="BASELINE configs[4]: full chain over a synthetic 3840x2160 frame stream + 3840x2160 backgrounds "
                        "-> 2160x2160; " + GENERATOR),
    "c1": dict(metric="images/sec symmetry x4 + crop_square, 1024x1024 RGB", batch=1000, e2e_frames=250,
               workload="BASELINE configs[0]: generate_symmetries (o, h, v, hv: one read, four writes) then "
                        "process_square_crop_around_bbox of every variant on 1024x1024x3 uniform-noise images"),
    "c2": dict(metric="images/sec rotation (expand + getbbox), 1920x1080 RGB", batch=512, e2e_frames=128,
               workload="BASELINE configs[1]: process_rotations, one rotation per image, angle ~ U(1, 359), RGB -> RGBA fused, "
                        "expanded canvas + alpha bbox; value = NEAREST (the reference's mode, bit-exact), value_bilinear = "
                        "BILINEAR (<= 1 LSB); 1920x1080x3 uniform-noise images"),
    "c3": dict(metric="images/sec overlay (LANCZOS resize + alpha paste + label), 1920x1080 backgrounds", batch=512, e2e_frames=256,
               workload="BASELINE configs[2]: paste_overlay_onto_background, 512x384 RGBA noise sprites with a radial alpha ramp "
                        "(0 / 255 plateaus), scale ~ U(.15, .30) of the background diagonal, onto 1920x1080x3 uniform-noise backgrounds"),
    "c4": dict(metric="images/sec HSV key + keep-largest-component, 512x512", batch=4096, e2e_frames=2048,
               workload="BASELINE configs[3]: process_images_with_color_masks (filter list of filtres_liste.py:186-190) then "
                        "keep_largest_component, fused (frame -> 1-bit mask -> largest component -> BGRA), on 512x512x3 synth.frame "
                        "images (blob + 1 % salt specks)"),
}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _shifts(n: int, seed: int, first: int, fh: int = FH, fw: int = FW):
    out = []
    for i in range(n):
        r = np.random.default_rng(seed * 7_000_003 + first + i)
        out.append((int(r.integers(-fh // 9, fh // 9 + 1)), int(r.integers(-fw * 5 // 32, fw * 5 // 32 + 1)), int(r.integers(0, fh)), int(r.integers(0, fw))))
    return out


def make_inputs(n: int, seed: int, first: int = 0, base: int = 8, fh: int = FH, fw: int = FW):
    """n frames + n backgrounds (numpy, host).  `base` distinct synthetic frames are generated with numpy
    (synth.frame, ~0.4 s each at 1080p) and multiplied by cyclic shifts so that every frame of the
    batch has different content and object position at negligible generation cost."""
    from image_processor_pipeline_b200 import synth
    bases = [synth.frame(seed, k, fh, fw) for k in range(base)]
    bgb = [synth.background(seed, k, fh, fw) for k in range(base)]
    frames = np.empty((n, fh, fw, 3), np.uint8)
    bgs = np.empty((n, fh, fw, 3), np.uint8)
    for i, (dy, dx, by, bx) in enumerate(_shifts(n, seed, first, fh, fw)):
        g = first + i
        frames[i] = np.roll(bases[g % base], (dy, dx), axis=(0, 1))
        bgs[i] = np.roll(bgb[g % base], (by, bx), axis=(0, 1))
    return frames, bgs


def make_inputs_device(n: int, seed: int, first: int, dev, base: int = 8, fh: int = FH, fw: int = FW, noise_frames: bool = False):
    """The same n frames + backgrounds as make_inputs, built on the device: only the `base` images are generated on
    the host; the cyclic shifts are gathers on the GPU (no n x 12 MB host copies per rank).
    noise_frames: frames are uniform noise (SURVEY.md §8d's generic generator) instead of synth.frame."""
    import torch
    from image_processor_pipeline_b200 import synth
    mk = (lambda k: synth.noise(seed + 13, k, (fh, fw, 3))) if noise_frames else (lambda k: synth.frame(seed, k, fh, fw))
    bases = torch.from_numpy(np.stack([mk(k) for k in range(base)])).to(dev)
    bgb = torch.from_numpy(np.stack([synth.background(seed, k, fh, fw) for k in range(base)])).to(dev)
    frames = torch.empty((n, fh, fw, 3), dtype=torch.uint8, device=dev)
    bgs = torch.empty((n, fh, fw, 3), dtype=torch.uint8, device=dev)
    sh = torch.tensor(_shifts(n, seed, first, fh, fw), dtype=torch.int64, device=dev)          # [n, 4] = dy, dx, by, bx
    which = (first + torch.arange(n, device=dev)) % base
    ys, xs = torch.arange(fh, device=dev), torch.arange(fw, device=dev)
    step = max(1, 64 * (FH * FW) // (fh * fw))
    for lo in range(0, n, step):                                                        # rolled[y, x] = src[(y - dy) % H, (x - dx) % W]
        sl = slice(lo, min(lo + step, n))
        w = which[sl, None, None]
        frames[sl] = bases[w, ((ys[None, :] - sh[sl, 0:1]) % fh)[:, :, None], ((xs[None, :] - sh[sl, 1:2]) % fw)[:, None, :]]
        bgs[sl] = bgb[w, ((ys[None, :] - sh[sl, 2:3]) % fh)[:, :, None], ((xs[None, :] - sh[sl, 3:4]) % fw)[:, None, :]]
    return frames, bgs


def noise_images_device(n: int, shape, seed: int, first: int, dev, base: int = 16):
    """n uniform-noise uint8 images [n, *shape] on the device: `base` images of SURVEY.md §8d's generator
    (default_rng(seed * 1_000_003 + i).integers(0, 256)) multiplied by cyclic row shifts."""
    import torch
    from image_processor_pipeline_b200 import synth
    bases = torch.from_numpy(np.stack([synth.noise(seed, k, shape) for k in range(base)])).to(dev)
    out = torch.empty((n,) + tuple(shape), dtype=torch.uint8, device=dev)
    for i in range(n):
        g = first + i
        out[i] = torch.roll(bases[g % base], shifts=(g * 37) % shape[0], dims=0)
    return out


def noise_images_host(n: int, shape, seed: int, first: int, base: int = 4):
    from image_processor_pipeline_b200 import synth
    bases = [synth.noise(seed, k, shape) for k in range(min(base, n))]
    return [np.roll(bases[(first + i) % len(bases)], ((first + i) * 37) % shape[0], axis=0) for i in range(n)]


def sprite_images(n: int, sh: int, sw: int, seed: int):
    """RGBA noise sprites with alpha = radial ramp with 0 / 255 plateaus."""
    r = np.random.default_rng(seed * 1_000_003 + 991)
    spr = r.integers(0, 256, (n, sh, sw, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:sh, 0:sw]
    spr[..., 3] = np.clip(255 * (1.2 - np.hypot((xx - sw / 2) / (sw / 2), (yy - sh / 2) / (sh / 2))), 0, 255).astype(np.uint8)
    return spr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm (oracle; every config)
C4_SHAPE, C1_SHAPE, C2_SHAPE, SPR = (512, 512), (1024, 1024, 3), (1080, 1920, 3), (384, 512)
_WORKER = {}


def _c1_box(i: int):
    r = np.random.default_rng(4242 + i)
    return np.array([[r.uniform(.3, .7), r.uniform(.3, .7), r.uniform(.05, .3), r.uniform(.05, .3)]])


def _cpu_worker_init(config, seed, count, worker_index, enhance=False):
    """Per-process set-up, outside any timed region: imports, single-threaded cv2, and this worker's own `count`
    synthetic inputs — seeded from the worker INDEX (reproducible; the process id plays no role)."""
    import cv2
    cv2.setNumThreads(1)
    from image_processor_pipeline_b200 import synth
    from oracle import chain as OC, libcalls as L
    first = 1000 * worker_index
    w = dict(config=config, OC=OC, L=L, filters=synth.REFERENCE_FILTERS, seed=seed + first, count=count, enhance=enhance)
    if config in ("chain", "c5"):
        c = CONFIGS[config]
        frames, bgs = make_inputs(count, seed, first, base=min(4, count), fh=c["fh"], fw=c["fw"])
        w.update(frames=list(frames), bgs=list(bgs))
    elif config == "c1":
        w.update(imgs=noise_images_host(count, C1_SHAPE, seed, first), boxes=[_c1_box(first + i) for i in range(count)])
    elif config == "c2":
        w.update(imgs=noise_images_host(count, C2_SHAPE, seed, first))
    elif config == "c3":
        w.update(bgs=noise_images_host(count, C2_SHAPE, seed, first), sprites=list(sprite_images(min(4, count), SPR[0], SPR[1], seed)))
    elif config == "c4":
        w.update(imgs=[synth.frame(seed, first + i, *C4_SHAPE) for i in range(min(8, count))])
    _WORKER.clear()
    _WORKER.update(w)


def _cpu_worker_run(step):
    """One pass of the oracle (the reference's cv2 / Pillow call sequence, in memory) over the worker's inputs.
    Returns (images produced, seconds of work)."""
    w = _WORKER
    L = w["L"]
    random.seed(w["seed"] + step)
    cfg = w["config"]
    t0 = time.perf_counter()
    if cfg in ("chain", "c5"):
        imgs, labels, _ = w["OC"].run_chain(w["frames"], w["bgs"], color_ranges_to_exclude_hsv=w["filters"], enhance=w.get("enhance", False), **CHAIN_KW)
        n = len(imgs)
    elif cfg == "c1":
        n = 0
        for img, box in zip(w["imgs"], w["boxes"]):
            syms = L.select_symmetries(None, None, True)
            for v in L.symmetries(img, syms):
                L.square_crop(v, np.array([0]), box)
            n += 1
    elif cfg == "c2":
        n = 0
        for img in w["imgs"]:
            L.rotate_and_fit(L.to_rgba(img), random.uniform(1.0, 359.0))
            n += 1
    elif cfg == "c3":
        n = 0
        for k, bg in enumerate(w["bgs"]):
            L.overlay_onto_background(w["sprites"][k % len(w["sprites"])], bg)
            n += 1
    else:
        n = 0
        for i in range(w["count"]):
            img = w["imgs"][i % len(w["imgs"])]
            L.keep_largest_component(L.color_masks_to_alpha(img, list(w["filters"])))
            n += 1
    return n, time.perf_counter() - t0


CPU_SAMPLE = {"chain": 16, "c5": 4, "c1": 32, "c2": 16, "c3": 32, "c4": 128}
REF_PER_WORKER = {"chain": 4, "c5": 1, "c1": 8, "c2": 4, "c3": 8, "c4": 64}


def cpu_rate_single(config: str, n_images: int, seed: int = 1, enhance: bool = False):
    """images/s of the oracle in this process, cv2 single-threaded."""
    _cpu_worker_init(config, seed, n_images, 0, enhance)
    if config not in ("chain", "c5"):
        _cpu_worker_run(0)                                    # warm the libraries (cheap configs only)
    done, dt = _cpu_worker_run(1)
    return done / dt, dt


def _ref_pool_init(config, per_worker, counter, enhance=False):
    with counter.get_lock():
        idx = counter.value
        counter.value += 1
    _cpu_worker_init(config, 100, per_worker, idx, enhance)


def run_reference_arm(args):
    """--impl reference: the reference's CPU path on all host cores.  ProcessPoolExecutor(cpu_count)
    is the reference's own workers=-1 mode (pipeline.py:88-90, :372); every worker owns
    `per_worker` inputs and one step = every worker running the oracle over its inputs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    per_worker = REF_PER_WORKER[args.config]
    per_step = per_worker * cores
    elapsed, produced = 0.0, 0
    counter = mp.Value("i", 0)
    with ProcessPoolExecutor(max_workers=cores, initializer=_ref_pool_init, initargs=(args.config, per_worker, counter, bool(args.enhance))) as ex:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = list(ex.map(_cpu_worker_run, [s] * cores, chunksize=1))
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                elapsed += dt
                produced += sum(r[0] for r in res)
    value = produced / elapsed
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["workload"] + ENHANCE_NOTE * bool(args.enhance), "name": args.config + "+enhance" * bool(args.enhance)},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} images per step ({per_worker} per worker) over a ProcessPoolExecutor({cores}); oracle = the reference's cv2/Pillow call sequence in memory (no file I/O), inputs generated outside the timed region, seeded per worker index"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm: workloads
class Workload:
    """One BASELINE config on one GPU: resident step, end-to-end step, roofline bookkeeping."""
    pipelined = False

    def extra(self):
        return {}


class ChainWorkload(Workload):
    pipelined = True

    def __init__(self, name, args, dev, rank):
        import torch
        from image_processor_pipeline_b200 import chain, synth
        self.torch, self.chain = torch, chain
        c = CONFIGS[name]
        self.name, self.args, self.dev, self.rank = name, args, dev, rank
        self.fh, self.fw = c["fh"], c["fw"]
        self.cs = min(self.fh, self.fw)
        self.B = args.batch or c["batch"]
        self.params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, enhance=bool(args.enhance), **CHAIN_KW)
        self.frames_d, self.bgs_d = make_inputs_device(self.B, 1, rank * self.B, dev, fh=self.fh, fw=self.fw)
        self.pipe = chain.ChainStream(self.params, stage_a_stream=args.stage_a_stream)
        self.alg_bytes_per_image = 2 * 3 * self.fw * self.fh + 3 * self.cs * self.cs       # frame + background + output
        self.alg = {"chain_mask_kernel": self.B * 3 * self.fw * self.fh,                   # frame read once
                    "chain_mask_seg_kernel": self.B * 3 * self.fw * self.fh,
                    "chain_paste_kernel": self.B * 2 * 3 * self.cs * self.cs}              # background crop read + output write
        self.l2_note = f"inputs per step {2 * self.B * 3 * self.fw * self.fh / 1e6:.0f} MB >> 126 MB L2 (no flush needed)"

    def step(self, s):
        random.seed(1000 * self.rank + s)
        return self.chain.run_chain(self.frames_d, self.bgs_d, self.params).images.shape[0]

    def run_pipelined(self, first, count, frames=None, bgs=None, pipe=None):
        """`count` steps through ChainStream.map: the four phases of consecutive steps are interleaved on one
        stream (finish(s-2), prepare(s-1), plan(s), submit(s+1) ...), so the device never waits for the host's
        planning or for the small read-backs.  Draw order = step after step (re-seeded per step)."""
        frames = self.frames_d if frames is None else frames
        bgs = self.bgs_d if bgs is None else bgs
        n_out = 0
        for res in (pipe or self.pipe).map(((frames, bgs) for _ in range(count)), random,
                                           on_batch=lambda i: random.seed(1000 * self.rank + first + i)):
            n_out += res.images.shape[0]
        return n_out

    def e2e_setup(self):
        torch = self.torch
        E = min(self.B, self.args.e2e_frames or CONFIGS[self.name]["e2e_frames"])
        self.E = E
        self.frames_pin = torch.empty((E, self.fh, self.fw, 3), dtype=torch.uint8).pin_memory()
        self.bgs_pin = torch.empty((E, self.fh, self.fw, 3), dtype=torch.uint8).pin_memory()
        self.frames_pin.copy_(self.frames_d[:E]); self.bgs_pin.copy_(self.bgs_d[:E])
        self.out_pin = torch.empty((E, self.cs, self.cs, 3), dtype=torch.uint8).pin_memory()
        self.last = {}

    def e2e_step(self, s):
        random.seed(1000 * self.rank + s)
        r = self.chain.run_chain_host(self.frames_pin, self.bgs_pin, self.params, self.dev, self.out_pin, chunk=self.args.e2e_chunk,
                                      background_windows=not self.args.no_bg_windows)
        self.last["h2d"], self.last["d2h"] = r.h2d_bytes, int(r.images.numel())
        return r.images.shape[0]

    def e2e_run(self, first, count):
        """`count` end-to-end steps as ONE stream through the host entry point (the pinned batch is streamed `count` times back to
        back: every step's frames cross PCIe again and every step's results come back, but the pipeline is not drained in between)."""
        random.seed(1000 * self.rank + first)
        r = self.chain.run_chain_host(self.frames_pin, self.bgs_pin, self.params, self.dev, self.out_pin, chunk=self.args.e2e_chunk,
                                      background_windows=not self.args.no_bg_windows, passes=count)
        self.last["h2d"], self.last["d2h"] = r.h2d_bytes // count, r.d2h_bytes // count
        return r.images_total

    def e2e_info(self):
        return {"pipelined_slice_frames": self.args.e2e_chunk, "frames_per_step": self.E,
                "steps_streamed_back_to_back": bool(self.args.e2e_stream),
                "background_upload": "whole images" if self.args.no_bg_windows else
                "square-crop windows only (planned offsets known before the copy; rows <= 4 KB: gathered from the pinned host buffer by a small kernel over PCIe, longer rows: cudaMemcpy2DAsync per item)"}

    def extra(self):
        """Content sensitivity: the same chain on UNIFORM-NOISE frames (SURVEY.md §8d's generic generator).  The key
        then keeps ~75 % of the pixels as one giant component: every pixel is classified, the sprite is the whole frame."""
        torch = self.torch
        if self.args.no_noise or self.name != "chain":
            return {}
        nb = min(self.B, 128)
        frames, bgs = make_inputs_device(nb, 2, self.rank * nb, self.dev, fh=self.fh, fw=self.fw, noise_frames=True)
        pipe = self.chain.ChainStream(self.params)
        self.run_pipelined(0, 3, frames, bgs, pipe)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = self.run_pipelined(3, 6, frames, bgs, pipe)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        from image_processor_pipeline_b200 import _native
        _native.timing_report()
        _native.timing_enable(True)                      # per-kernel device time of one more (untimed) noise step
        self.run_pipelined(9, 1, frames, bgs, pipe)
        rep = _native.timing_report()
        _native.timing_enable(False)
        del frames, bgs
        return {"value_noise_frames": n / (ms * 1e-3),
                "noise_frames_kernels_ms_per_step": {k: round(v[1], 4) for k, v in sorted(rep.items())},
                "noise_frames_note": f"same chain, frames = uniform noise (worst case: every pixel classified, ~75 % of the frame is one component and becomes the sprite), {nb} frames x 6 steps, this rank only"}


class OpsWorkload(Workload):
    """c1..c4: one transform (or two chained steps) mapped over a resident batch through the batched operators."""

    def __init__(self, name, args, dev, rank):
        import torch
        from image_processor_pipeline_b200 import batch, ops, synth
        self.torch, self.batch, self.ops, self.synth = torch, batch, ops, synth
        self.name, self.args, self.dev, self.rank = name, args, dev, rank
        self.B = B = args.batch or CONFIGS[name]["batch"]
        first = rank * B
        if name == "c1":
            self.imgs = noise_images_device(B, C1_SHAPE, 1, first, dev)
            H, W, _ = C1_SHAPE
            self.alg_bytes_per_image = 5 * H * W * 3 + 4 * 2 * H * W * 3                    # symmetries 1 read + 4 writes; 4 crops read + write
            self.alg = {"rows_kernel": None}
        elif name == "c2":
            self.imgs = noise_images_device(B, C2_SHAPE, 1, first, dev)
            H, W, _ = C2_SHAPE
            self.alg_bytes_per_image = 3 * W * H + 4 * (W * H + (W * W + H * H) / np.pi)    # SURVEY.md §8d, RGB -> RGBA fused, mean over the angle
            self.alg = {"rotate_batch_nearest_kernel": B * self.alg_bytes_per_image, "rotate_batch_bilinear_kernel": B * self.alg_bytes_per_image}
        elif name == "c3":
            self.imgs = noise_images_device(B, C2_SHAPE, 1, first, dev)                     # the backgrounds
            self.sprites = torch.from_numpy(sprite_images(64, SPR[0], SPR[1], 1)).to(dev)
            H, W, _ = C2_SHAPE
            self.alg_bytes_per_image = 2 * 3 * W * H + 4 * SPR[0] * SPR[1] + 4 * 0.047 * W * H   # bg read + composite write + sprite source + resized sprite (mean 4.7 % of the bg)
            self.alg = {"chain_paste_kernel": B * 2 * 3 * W * H}
        else:
            base = torch.from_numpy(np.stack([synth.frame(1, first + k, *C4_SHAPE) for k in range(32)])).to(dev)
            self.imgs = torch.empty((B,) + C4_SHAPE + (3,), dtype=torch.uint8, device=dev)
            for i in range(B):
                self.imgs[i] = torch.roll(base[i % 32], shifts=((i * 53) % 512, (i * 97) % 512), dims=(0, 1))
            H, W = C4_SHAPE
            self.alg_bytes_per_image = 3 * H * W + 4 * H * W                                 # BGR read + BGRA write (HSV key); the CCL works on bits
            self.alg = {"chain_mask_seg_kernel": B * 3 * H * W, "chain_mask_kernel": B * 3 * H * W, "bgra_from_keep_kernel": B * 7 * H * W}
        self.l2_note = f"inputs per step {self.imgs.numel() / 1e6:.0f} MB >> 126 MB L2 (no flush needed)"

    # the work on device tensors (also used by the end-to-end leg on uploaded slices); returns (n images, result tensors)
    def work(self, x, s, mode=None):
        ops, batch = self.ops, self.batch
        if self.name == "c1":
            from image_processor_pipeline_b200 import _native
            n, H, W, c = x.shape
            outs = [self.torch.empty_like(x) for _ in range(4)]
            for _ in range(n):
                random.sample(["o", "h", "v", "hv"], 4)                       # symmetry.py:122: the draw fixes the output order only
            _native.check(_native.load().ipp_symmetries_u8(x.data_ptr(), W * c, W, H, c, *[o.data_ptr() for o in outs], W * c, n, H * W * c, H * W * c,
                                                           self.torch.cuda.current_stream().cuda_stream), "ipp_symmetries_u8")
            cs = min(H, W)
            res = []
            for o in outs:                                                    # crop_square.py:172-196 per variant: two randint per image
                xy = np.array([[random.randint(0, W - cs), random.randint(0, H - cs)] for _ in range(n)], dtype=np.int32)
                res.append(batch.crop_batch(o, xy, cs, cs))
            return n, res
        if self.name == "c2":
            r = batch.rotate_batch(x, random, 1.0, 359.0, resample=mode or "nearest")
            return x.shape[0], [r.pool, r._bbox_d]
        if self.name == "c3":
            r = batch.overlay_batch(self.sprites, x, x.shape[0], random)
            return r.images.shape[0], [r.images]
        cleaned, info = batch.key_and_keep_largest(x, self.synth.REFERENCE_FILTERS)
        return x.shape[0], [cleaned]

    def step(self, s, mode=None):
        random.seed(1000 * self.rank + s)
        return self.work(self.imgs, s, mode)[0]

    def e2e_setup(self):
        torch = self.torch
        E = min(self.B, self.args.e2e_frames or CONFIGS[self.name]["e2e_frames"])
        self.E = E
        self.in_pin = torch.empty((E,) + tuple(self.imgs.shape[1:]), dtype=torch.uint8).pin_memory()
        self.in_pin.copy_(self.imgs[:E])
        self.chunk = max(1, E // 8)
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(2)]
        self.out_pin = {}
        self.last = {}

    def e2e_step(self, s):
        """pinned host inputs -> H2D -> batched operators -> D2H of every result, slices alternating between two streams
        (copies of one slice overlap the kernels of the other)."""
        torch = self.torch
        random.seed(1000 * self.rank + s)
        n_out, h2d, d2h = 0, 0, 0
        for ci, lo in enumerate(range(0, self.E, self.chunk)):
            st = self.streams[ci % 2]
            with torch.cuda.stream(st):
                xin = self.in_pin[lo:lo + self.chunk].to(self.dev, non_blocking=True)
                h2d += xin.numel()
                n, outs = self.work(xin, s)
                n_out += n
                for k, o in enumerate(outs):
                    key = (ci, k)
                    buf = self.out_pin.get(key)
                    if buf is None or buf.numel() < o.numel():
                        buf = self.out_pin[key] = torch.empty(o.numel(), dtype=o.dtype).pin_memory()
                    buf[:o.numel()].copy_(o.reshape(-1), non_blocking=True)
                    d2h += o.numel() * o.element_size()
        for st in self.streams:
            st.synchronize()
        self.last["h2d"], self.last["d2h"] = h2d, d2h
        return n_out

    def e2e_info(self):
        return {"frames_per_step": self.E, "slice_frames": self.chunk, "streams": 2}

    def extra(self):
        if self.name != "c2":
            return {}
        torch = self.torch
        for s in range(2):
            self.step(s, "bilinear")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = sum(self.step(10 + s, "bilinear") for s in range(5))
        e1.record()
        torch.cuda.synchronize()
        return {"value_bilinear": n / (e0.elapsed_time(e1) * 1e-3)}


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: enough for >= 1 s of timed work)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="chain", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default per config)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-noise", action="store_true", help="chain: skip the uniform-noise-frames line")
    ap.add_argument("--enhance", action="store_true", help="chain / c5: add tranfo.py's enhance_image (pairing 'sample') as the last step")
    ap.add_argument("--no-pipeline", action="store_true", help="chain: one run_chain call per step (host waits inside the step)")
    ap.add_argument("--stage-a-stream", default="caller", choices=["side", "caller"], help="stage A of new batches on a second stream")
    ap.add_argument("--no-bg-windows", action="store_true", help="e2e: upload whole backgrounds instead of the planned crop windows")
    ap.add_argument("--e2e-frames", type=int, default=0, help="images of the batch used by the end-to-end leg (default per config)")
    ap.add_argument("--e2e-stream", action="store_true", help="chain e2e: all steps as ONE stream through run_chain_host (passes=steps) instead of one call per step")
    ap.add_argument("--e2e-chunk", type=int, default=16, help="frames per pipelined slice of the chain's host-buffer entry point")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"chain": 160, "c5": 60, "c1": 100, "c2": 100, "c3": 120, "c4": 100}[args.config] if args.impl == "b200" else 5
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from image_processor_pipeline_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the one JSON line only: NCCL prints its version banner there when the communicator comes up, so the
        # communicator is created (init + one barrier) with fd 1 pointing at stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    _native.load()
    cfg = CONFIGS[args.config]
    wl = (ChainWorkload if args.config in ("chain", "c5") else OpsWorkload)(args.config, args, dev, rank)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(ms, n_out):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            c = torch.tensor([n_out], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            n_out = int(c.item())
        return ms, n_out

    def timed(fn, steps, warmup):
        for s in range(warmup):
            fn(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_out = 0
        l0 = _native.launch_count()
        e0.record()
        for s in range(steps):
            n_out += fn(warmup + s)
        e1.record()
        barrier()
        ms, n_out = reduce(e0.elapsed_time(e1), n_out)
        return ms, n_out, _native.launch_count() - l0

    def timed_pipelined(steps, warmup):
        wl.run_pipelined(0, warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _native.launch_count()
        e0.record()
        n_out = wl.run_pipelined(warmup, steps)
        e1.record()
        barrier()
        ms, n_out = reduce(e0.elapsed_time(e1), n_out)
        return ms, n_out, _native.launch_count() - l0

    with ClockSampler(local) as clk:
        if wl.pipelined and not args.no_pipeline:
            ms, n_out, launches = timed_pipelined(args.steps, args.warmup)
        else:
            ms, n_out, launches = timed(wl.step, args.steps, args.warmup)
    clocks = clk.summary()
    value = n_out / (ms * 1e-3)

    # per-kernel device time of a few of the same steps (events around every launch) -> roofline
    _native.timing_enable(True)
    ksteps = min(args.steps, 10)
    for s in range(ksteps):
        wl.step(args.warmup + s)
    torch.cuda.synchronize()
    rep = _native.timing_report()
    _native.timing_enable(False)
    peak, peak_src = measured_peak()
    kern = {k: {"launches": c, "total_ms": round(t, 4), "avg_ms": round(t / c, 5)} for k, (c, t) in rep.items()}
    total_kernel_ms = sum(t for _, t in rep.values())
    alg = dict(wl.alg)
    if "rows_kernel" in alg:                                            # c1: five launches of the same kernel per step (symmetries + 4 crops)
        alg["rows_kernel"] = wl.B * wl.alg_bytes_per_image / 5
    alg = {k: v for k, v in alg.items() if k in rep}
    dom = max(alg, key=lambda k: rep[k][1])
    dcnt, dms = rep[dom]
    achieved = alg[dom] / (dms / dcnt * 1e-3) / 1e9
    traffic, traffic_src = None, None                                  # dram bytes per launch from the committed ncu --set full capture
    tp = ROOT / "profiles" / "traffic.json"
    if tp.is_file() and args.config in ("chain", "c5", "c2", "c4"):      # per-frame figures of 1080p / C2 / C4 captures (c5: 4K frames = 4 x the bytes)
        rec = json.loads(tp.read_text()).get(dom)
        if rec:
            traffic = rec["dram_bytes_per_frame"] * wl.B * (4 if args.config == "c5" else 1)
            traffic_src = rec["source"] + (" — scaled x4 from the 1080p capture to 3840x2160 frames" if args.config == "c5" else "")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": dms / dcnt,
                "share_of_kernel_time": dms / total_kernel_ms if total_kernel_ms else None}
    alg_gbs = wl.alg_bytes_per_image * (n_out / world / args.steps) / (ms / args.steps * 1e-3) / 1e9

    extra = wl.extra()

    # end to end through the host API: pinned host inputs, H2D + work + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        wl.e2e_setup()
        esteps = {"chain": 6, "c5": 6}.get(args.config, 10)
        if hasattr(wl, "e2e_run") and args.e2e_stream:
            wl.e2e_run(0, 2)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            en = wl.e2e_run(2, esteps)
            e1.record()
            barrier()
            ems, en = reduce(e0.elapsed_time(e1), en)
        else:
            ems, en, _ = timed(wl.e2e_step, esteps, 2)
        e2e = {"value": en / (ems * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(wl.last["h2d"]), "d2h_bytes_per_step": int(wl.last["d2h"])}
        e2e.update(wl.e2e_info())

    if rank == 0:
        cpu = None
        if world == 1:
            sample = args.cpu_sample or CPU_SAMPLE[args.config]
            rate, dt = cpu_rate_single(args.config, sample, enhance=bool(args.enhance))
            cpu = {"value": rate, "unit": "images/s", "cores": 1, "kind": "port",
                   "sample": f"{sample} images of the same workload, oracle (the reference's cv2/Pillow call sequence in memory, cv2.setNumThreads(1)), {dt:.1f} s",
                   "host_cores_available": os.cpu_count()}
        line = {"metric": cfg["metric"], "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": cfg["workload"] + ENHANCE_NOTE * bool(args.enhance), "name": args.config + "+enhance" * bool(args.enhance)},
                "run": {"images_per_gpu_per_step": wl.B, "software_pipelined_steps": bool(wl.pipelined and not args.no_pipeline), "l2": wl.l2_note,
                        "host_work_in_step": "native planner (Python-compatible MT19937 draws) + small pinned read-backs; the chain interleaves the phases of consecutive steps on one stream" if wl.pipelined else "draws from Python's random in the reference's order (native planner for rotations / overlays)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "algorithmic_gbs": alg_gbs, "frac_of_peak": alg_gbs / peak,
                "algorithmic_bytes_per_image": wl.alg_bytes_per_image,
                "kernels": kern, "cpu_baseline": cpu}
        if args.config == "chain":
            line["chain_algorithmic_gbs"], line["chain_frac_of_peak"] = alg_gbs, alg_gbs / peak
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
