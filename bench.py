#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native transform chain.

Metric (BASELINE.json): images/sec, 1080p RGB full chain, at 1/2/4/8 B200; % of HBM roofline.
A "step" = one pass of the full chain (border crop -> HSV key -> largest component + fit ->
rotate + fit -> flip -> LANCZOS overlay -> square crop, with YOLO labels) over one batch of
synthetic 1920x1080 frames and backgrounds per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

* value      frames/s, inputs resident in HBM, CUDA-event timed, max over ranks, all N GPUs.
* e2e        same metric through the public host API (chain.run_chain_host): pinned host
             frames/backgrounds -> H2D -> chain -> D2H of the outputs inside the timed region.
* roofline   dominant kernel: algorithmic bytes per launch / its mean CUDA-event duration
             (events on the launching stream around every launch, ipp_timing_* hooks).
* cpu_baseline  the oracle chain (the reference's own cv2/Pillow call sequence, in memory)
             on this box's host cores, bounded sample.  Rank 0, N=1 only.
--impl reference times that CPU path with all host cores (ProcessPoolExecutor, the
reference's workers=-1 mode, pipeline.py:88-90,372).
"""
from __future__ import annotations

import argparse
import collections
import contextlib
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
ALG_BYTES_PER_IMAGE = 3 * FW * FH + 3 * FW * FH + 3 * 1080 * 1080      # frame + background + output
WORKLOAD = "full chain (border-crop + HSV key + largest component + rotate + flip + LANCZOS overlay + square crop + labels), 1920x1080 RGB frame + 1920x1080 background -> 1080x1080"
CHAIN_KW = dict(num_rotations=1, include_original=False, pool=["o", "h", "v", "hv"], choose_random=1,
                sym_include_original=False)


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _shifts(n: int, seed: int, first: int):
    out = []
    for i in range(n):
        r = np.random.default_rng(seed * 7_000_003 + first + i)
        out.append((int(r.integers(-120, 121)), int(r.integers(-300, 301)), int(r.integers(0, FH)), int(r.integers(0, FW))))
    return out


def make_inputs(n: int, seed: int, first: int = 0, base: int = 8):
    """n frames + n backgrounds (numpy, host).  `base` distinct synthetic frames are generated with numpy
    (synth.frame, ~0.4 s each) and multiplied by cyclic shifts so that every frame of the
    batch has different content and object position at negligible generation cost."""
    from image_processor_pipeline_b200 import synth
    bases = [synth.frame(seed, k, FH, FW) for k in range(base)]
    bgb = [synth.background(seed, k, FH, FW) for k in range(base)]
    frames = np.empty((n, FH, FW, 3), np.uint8)
    bgs = np.empty((n, FH, FW, 3), np.uint8)
    for i, (dy, dx, by, bx) in enumerate(_shifts(n, seed, first)):
        g = first + i
        frames[i] = np.roll(bases[g % base], (dy, dx), axis=(0, 1))
        bgs[i] = np.roll(bgb[g % base], (by, bx), axis=(0, 1))
    return frames, bgs


def make_inputs_device(n: int, seed: int, first: int, dev, base: int = 8):
    """The same n frames + backgrounds as make_inputs, built on the device: only the `base` images are generated on
    the host; the cyclic shifts are torch.roll on the GPU (no n x 12 MB host copies per rank)."""
    import torch
    from image_processor_pipeline_b200 import synth
    bases = torch.from_numpy(np.stack([synth.frame(seed, k, FH, FW) for k in range(base)])).to(dev)
    bgb = torch.from_numpy(np.stack([synth.background(seed, k, FH, FW) for k in range(base)])).to(dev)
    frames = torch.empty((n, FH, FW, 3), dtype=torch.uint8, device=dev)
    bgs = torch.empty((n, FH, FW, 3), dtype=torch.uint8, device=dev)
    sh = torch.tensor(_shifts(n, seed, first), dtype=torch.int64, device=dev)          # [n, 4] = dy, dx, by, bx
    which = (first + torch.arange(n, device=dev)) % base
    ys, xs = torch.arange(FH, device=dev), torch.arange(FW, device=dev)
    for lo in range(0, n, 64):                                                          # rolled[y, x] = src[(y - dy) % H, (x - dx) % W]: one gather per 64 images
        sl = slice(lo, min(lo + 64, n))
        w = which[sl, None, None]
        frames[sl] = bases[w, ((ys[None, :] - sh[sl, 0:1]) % FH)[:, :, None], ((xs[None, :] - sh[sl, 1:2]) % FW)[:, None, :]]
        bgs[sl] = bgb[w, ((ys[None, :] - sh[sl, 2:3]) % FH)[:, :, None], ((xs[None, :] - sh[sl, 3:4]) % FW)[:, None, :]]
    return frames, bgs


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


# ------------------------------------------------------------------------------------ CPU arm
_WORKER = {}


def _cpu_worker_init(seed, count):
    """Per-process set-up, outside any timed region: imports, single-threaded cv2, and this
    worker's own `count` synthetic frames/backgrounds."""
    import cv2
    cv2.setNumThreads(1)
    from image_processor_pipeline_b200 import synth
    from oracle import chain as OC
    first = (os.getpid() * 7919) % 100_000
    frames, bgs = make_inputs(count, seed, first, base=min(4, count))
    _WORKER.update(frames=list(frames), bgs=list(bgs), OC=OC, filters=synth.REFERENCE_FILTERS, seed=seed + first)


def _cpu_worker_run(step):
    """One pass of the oracle chain (the reference's cv2/Pillow call sequence, in memory) over the
    worker's frames.  Returns (images produced, seconds of chain work)."""
    w = _WORKER
    random.seed(w["seed"] + step)
    t0 = time.perf_counter()
    imgs, labels, _ = w["OC"].run_chain(w["frames"], w["bgs"], color_ranges_to_exclude_hsv=w["filters"], **CHAIN_KW)
    return len(imgs), time.perf_counter() - t0


def cpu_chain_rate_single(n_images: int, seed: int = 1):
    """images/s of the oracle chain in this process, cv2 single-threaded."""
    _cpu_worker_init(seed, n_images)
    _cpu_worker_run(0) if n_images <= 4 else None            # warm the libraries on tiny samples only
    done, dt = _cpu_worker_run(1)
    return done / dt, dt


def run_reference_arm(args):
    """--impl reference: the reference's CPU path on all host cores.  ProcessPoolExecutor(cpu_count)
    is the reference's own workers=-1 mode (pipeline.py:88-90, :372); every worker owns
    `per_worker` frames and one step = every worker running the oracle chain over its frames."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    per_worker = 4
    per_step = per_worker * cores
    elapsed, produced = 0.0, 0
    with ProcessPoolExecutor(max_workers=cores, initializer=_cpu_worker_init, initargs=(100, per_worker)) as ex:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = list(ex.map(_cpu_worker_run, [s] * cores, chunksize=1))
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                elapsed += dt
                produced += sum(r[0] for r in res)
    value = produced / elapsed
    line = {"impl": "reference", "metric": "images/sec 1080p RGB full chain", "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} frames per step ({per_worker} per worker) over a ProcessPoolExecutor({cores}); oracle chain = the reference's cv2/Pillow call sequence in memory (no file I/O), inputs generated outside the timed region"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="do not overlap stage A of step s+1 with the tail of step s")
    ap.add_argument("--stage-a-stream", default="caller", choices=["side", "caller"], help="stage A of new batches on a second stream")
    ap.add_argument("--no-bg-windows", action="store_true", help="e2e: upload whole backgrounds instead of the planned crop windows")
    ap.add_argument("--e2e-frames", type=int, default=512, help="frames of the batch used by the end-to-end leg (0 = all; 512 = 8 GB of pinned host memory per rank)")
    ap.add_argument("--e2e-chunk", type=int, default=16, help="frames per pipelined slice of the host-buffer entry point")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from image_processor_pipeline_b200 import _native, chain, synth

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
    B = args.batch
    params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **CHAIN_KW)

    # each rank owns a contiguous slice of the global work list (weak scaling: B frames per GPU)
    frames_d, bgs_d = make_inputs_device(B, seed=1, first=rank * B, dev=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(s):
        random.seed(1000 * rank + s)
        return chain.run_chain(frames_d, bgs_d, params)

    pipe = chain.ChainStream(params, stage_a_stream=args.stage_a_stream)

    def run_pipelined(first, count):
        """`count` steps through ChainStream.map: the four phases of consecutive steps are interleaved on one
        stream (finish(s-3), prepare(s-2), plan(s-1), submit(s)), so the device never waits for the host's
        planning or for the small read-backs.  Draw order = step after step (re-seeded per step)."""
        n_out = 0
        for res in pipe.map(((frames_d, bgs_d) for _ in range(count)), random,
                            on_batch=lambda i: random.seed(1000 * rank + first + i)):
            n_out += res.images.shape[0]
        return n_out

    def timed(fn, steps, warmup):
        for s in range(warmup):
            fn(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_out = 0
        l0 = _native.launch_count()
        e0.record()
        for s in range(steps):
            n_out += fn(warmup + s).images.shape[0]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _native.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            c = torch.tensor([n_out], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            n_out = int(c.item())
        return ms, n_out, launches

    def timed_pipelined(steps, warmup):
        run_pipelined(0, warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _native.launch_count()
        e0.record()
        n_out = run_pipelined(warmup, steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _native.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            c = torch.tensor([n_out], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            n_out = int(c.item())
        return ms, n_out, launches

    # stages B2/C (the caller's stream) get a higher priority than stage A of the next batch (side stream):
    # the big persistent mask CTAs then fill SMs as the finishing batch drains instead of delaying it
    with ClockSampler(local) as clk:
        if args.no_pipeline:
            ms, n_out, launches = timed(step_resident, args.steps, args.warmup)
        else:
            ms, n_out, launches = timed_pipelined(args.steps, args.warmup)
    clocks = clk.summary()
    value = n_out / (ms * 1e-3)

    # per-kernel device time of the same steps (events around every launch) -> roofline
    _native.timing_enable(True)
    for s in range(args.steps):
        step_resident(args.warmup + s)
    torch.cuda.synchronize()
    rep = _native.timing_report()
    _native.timing_enable(False)
    alg = {"chain_mask_kernel": B * 3 * FW * FH,                       # frame read once (generic mask kernel)
           "chain_mask_seg_kernel": B * 3 * FW * FH,                   # frame read once (table-driven per-warp segment kernel)
           "chain_paste_kernel": B * 2 * 3 * 1080 * 1080}              # background crop read + output write
    peak, peak_src = measured_peak()
    kern = {k: {"launches": c, "total_ms": round(t, 4), "avg_ms": round(t / c, 5)} for k, (c, t) in rep.items()}
    total_kernel_ms = sum(t for _, t in rep.values())
    dom = max(alg, key=lambda k: rep.get(k, (1, 0.0))[1])
    dcnt, dms = rep.get(dom, (1, float("nan")))
    achieved = alg[dom] / (dms / dcnt * 1e-3) / 1e9
    traffic, traffic_src = None, None                                  # dram bytes per launch from the committed ncu --set full capture
    tp = ROOT / "profiles" / "traffic.json"
    if tp.is_file():
        rec = json.loads(tp.read_text()).get(dom)
        if rec:
            traffic = rec["dram_bytes_per_frame"] * B
            traffic_src = rec["source"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": dms / dcnt,
                "share_of_kernel_time": dms / total_kernel_ms if total_kernel_ms else None}
    chain_gbs = ALG_BYTES_PER_IMAGE * (n_out / world / args.steps) / (ms / args.steps * 1e-3) / 1e9

    # end to end through the host API: pinned host inputs, H2D + chain + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        # host buffers of the end-to-end leg: the same batch in pinned memory (B = 512: 8 GB per rank; the first slice's
        # upload is not overlapped, so short batches under-report the PCIe-bound steady state: 4.80 k at 256, 5.05 k at 512)
        E = min(B, args.e2e_frames) if args.e2e_frames else B
        frames_pin = torch.empty((E, FH, FW, 3), dtype=torch.uint8).pin_memory()
        bgs_pin = torch.empty((E, FH, FW, 3), dtype=torch.uint8).pin_memory()
        frames_pin.copy_(frames_d[:E]); bgs_pin.copy_(bgs_d[:E])
        out_pin = torch.empty((E, 1080, 1080, 3), dtype=torch.uint8).pin_memory()

        last = {}

        def step_e2e(s):
            random.seed(1000 * rank + s)
            r = chain.run_chain_host(frames_pin, bgs_pin, params, dev, out_pin, chunk=args.e2e_chunk,
                                     background_windows=not args.no_bg_windows)
            last["h2d"], last["d2h"] = r.h2d_bytes, int(r.images.numel())
            return r

        ems, en, _ = timed(step_e2e, max(2, args.steps // 2), 1)
        e2e = {"value": en / (ems * 1e-3), "unit": "images/s", "pipelined_slice_frames": args.e2e_chunk,
               "frames_per_step": E, "h2d_bytes_per_step": int(last["h2d"]), "d2h_bytes_per_step": int(last["d2h"]),
               "background_upload": "whole images" if args.no_bg_windows else
               "square-crop windows only (planned offsets known before the copy; cudaMemcpy2DAsync per item)"}

    if rank == 0:
        cpu = None
        if world == 1:
            rate, dt = cpu_chain_rate_single(args.cpu_sample)
            cpu = {"value": rate, "unit": "images/s", "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_sample} frames of the same workload, oracle chain (the reference's cv2/Pillow call sequence in memory, cv2.setNumThreads(1)), {dt:.1f} s",
                   "host_cores_available": os.cpu_count()}
        line = {"metric": "images/sec 1080p RGB full chain", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": B, "software_pipelined_steps": not args.no_pipeline,
                           "l2": f"inputs per step {2 * B * 3 * FW * FH / 1e6:.0f} MB >> 126 MB L2 (no flush needed)",
                           "host_work_in_step": "C++ planner (Python-compatible MT19937 draws) + 3 small device->host read-backs per batch (RNG bounds depend on sprite sizes; table-ambiguity count)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "chain_algorithmic_gbs": chain_gbs, "chain_frac_of_peak": chain_gbs / peak,
                "kernels": kern, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
