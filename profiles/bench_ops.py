"""Per-operator throughput at the sizes of BASELINE.json's configs 1-4 (the stand-alone C-ABI entry points behind
transforms/*; the headline chain is bench.py).  Device-resident inputs, CUDA events, algorithmic bytes of
SURVEY.md §8d.  python profiles/bench_ops.py"""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from image_processor_pipeline_b200 import _native, ops, synth

dev = torch.device('cuda:0')
PEAK = 6544.7
try:
    PEAK = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception:
    pass


def noise(shape, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randint(0, 256, shape, dtype=torch.uint8, device=dev, generator=g)


def timed(fn, inputs, reps=3):
    """fn over a pool of distinct inputs (>> L2), best of `reps` passes; returns ms per call."""
    for x in inputs[:4]:
        fn(x)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for x in inputs:
            fn(x)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / len(inputs))
    # device time of the library's kernels alone (events around every launch)
    _native.timing_enable(True)
    for x in inputs:
        fn(x)
    torch.cuda.synchronize()
    rep = _native.timing_report()
    _native.timing_enable(False)
    timed.kernel_ms = sum(t for _, t in rep.values()) / len(inputs)
    timed.kernels = "+".join(sorted(rep))
    return best


rows = []


def report(name, ms, alg_bytes, scale=1.0):
    """ms: wall (CUDA events around the Python calls, host overhead included); timed.kernel_ms: kernels only."""
    kms = timed.kernel_ms * scale
    gbs, kgbs = alg_bytes / (ms * 1e-3) / 1e9, alg_bytes / (kms * 1e-3) / 1e9
    rows.append(dict(op=name, ms_per_call=round(ms, 4), per_s=round(1e3 / ms, 1), algorithmic_MB=round(alg_bytes / 1e6, 2), GBs_call=round(gbs, 1),
                     kernel_ms=round(kms, 4), GBs_kernels=round(kgbs, 1), frac_of_measured_peak_kernels=round(kgbs / PEAK, 3), kernels=timed.kernels))
    print(f"{name:56s} call {ms:7.4f} ms {gbs:7.1f} GB/s | kernels {kms:7.4f} ms {kgbs:7.1f} GB/s = {100 * kgbs / PEAK:5.1f} % of {PEAK:.0f}  [{timed.kernels}]")


# C1: symmetries (o, h, v, hv) + square crop, 1024 x 1024 x 3
pool = [noise((1024, 1024, 3), i) for i in range(96)]
report("C1 symmetries o+h+v+hv 1024x1024x3 (1 read, 4 writes)", timed(lambda x: ops.symmetries(x, ["o", "h", "v", "hv"]), pool), 5 * 1024 * 1024 * 3)
pool = [noise((1080, 1920, 3), i) for i in range(64)]
report("C1' square crop 1920x1080 -> 1080x1080", timed(lambda x: ops.flip_crop(x, 419, 0, 1080, 1080), pool), 2 * 1080 * 1080 * 3)
# C2: rotation (nearest, expand) + bbox, 1920 x 1080 RGBA
pool = [ops.to_rgba(noise((1080, 1920, 3), 100 + i)) for i in range(48)]
angles = np.random.default_rng(0).uniform(1, 359, len(pool))
it = iter(range(10 ** 9))
ms = timed(lambda x: ops.rotate_expand(x, float(angles[next(it) % len(pool)]), with_bbox=True), pool)
report("C2 rotate nearest expand + bbox 1920x1080 RGBA", ms, 4 * 1920 * 1080 + 4 * 3.618e6)
# C3: overlay (LANCZOS resize of a 512x384 sprite + paste) on 1920 x 1080
bgs = [noise((1080, 1920, 3), 200 + i) for i in range(64)]
spr = noise((384, 512, 4), 7)
small = ops.resize_lanczos(spr, (396, 297))
report("C3 paste 396x297 RGBA sprite on 1920x1080", timed(lambda b: ops.paste(b, small, (700, 300)), bgs), 2 * 3 * 1920 * 1080 + 4 * 396 * 297)
report("C3 resize LANCZOS 512x384 -> 396x297 RGBA (scalar op)", timed(lambda b: ops.resize_lanczos(spr, (396, 297)), bgs[:16]), 4 * 512 * 384 + 4 * 396 * 297)
# C4: HSV filter list -> BGRA, keep largest component, 512 x 512, batches of 256
frames = torch.from_numpy(np.stack([synth.frame(3, i, 512, 512) for i in range(8)])).to(dev).repeat(32, 1, 1, 1)
batches = [torch.roll(frames, (17 * k, 29 * k), (1, 2)) for k in range(6)]
n = frames.shape[0]
ms = timed(lambda x: ops.color_masks_to_alpha(x, synth.REFERENCE_FILTERS), batches)
report(f"C4 HSV filter list -> BGRA 512x512 (batch of {n})", ms / n, 7 * 512 * 512, 1.0 / n)
keyed = [ops.color_masks_to_alpha(x, synth.REFERENCE_FILTERS) for x in batches]
ms = timed(lambda x: ops.keep_largest_component(x), keyed)
report(f"C4 keep largest component 512x512 BGRA (batch of {n})", ms / n, 2 * 4 * 512 * 512, 1.0 / n)
json.dump(rows, open("gpurun_out/bench_ops.json", "w"), indent=1)
