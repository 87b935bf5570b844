"""Device timeline of the software-pipelined chain (ChainStream): start / end of every kernel launch per stream
(ipp_timing_timeline: CUDA events around each launch; all streams share the device clock).  Shows which kernels
actually run concurrently and how much of a step the GPU spends with one / two / no kernel resident.

    python profiles/timeline.py [--batch 1024] [--steps 4] [--out gpurun_out/timeline.json]
"""
import argparse
import collections
import json
import random
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from image_processor_pipeline_b200 import _native, chain, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--out", default="gpurun_out/timeline.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    _native.load()
    params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **bench.CHAIN_KW)
    frames, bgs = bench.make_inputs_device(args.batch, 1, 0, dev)
    pipe = chain.ChainStream(params)

    def run(count, first):
        for _ in pipe.map(((frames, bgs) for _ in range(count)), random, on_batch=lambda i: random.seed(first + i)):
            pass

    run(5, 0)
    torch.cuda.synchronize()
    _native.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(args.steps, 10)
    e1.record()
    torch.cuda.synchronize()
    tl = _native.timing_timeline()
    _native.timing_enable(False)
    total = e0.elapsed_time(e1)
    # coverage: time with >= 1 and >= 2 kernels in flight
    ev = sorted([(t0, 1) for _, _, t0, _ in tl] + [(t1, -1) for _, _, _, t1 in tl])
    depth, last, busy1, busy2 = 0, ev[0][0], 0.0, 0.0
    for t, d in ev:
        if depth >= 1:
            busy1 += t - last
        if depth >= 2:
            busy2 += t - last
        depth += d
        last = t
    span = max(t1 for *_, t1 in tl) - min(t0 for _, _, t0, _ in tl)
    per = collections.defaultdict(float)
    for name, _, t0, t1 in tl:
        per[name] += t1 - t0
    summary = {"batch": args.batch, "steps": args.steps, "wall_ms": total, "span_ms": span, "busy_ge1_ms": busy1, "busy_ge2_ms": busy2,
               "sum_kernel_ms": sum(per.values()), "per_kernel_ms_per_step": {k: round(v / args.steps, 4) for k, v in sorted(per.items())}}
    print(json.dumps(summary))
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps({"summary": summary, "launches": tl}))
    # compact text view of the middle step
    lo = span * (args.steps // 2) / args.steps
    hi = span * (args.steps // 2 + 1) / args.steps
    t_base = min(t0 for _, _, t0, _ in tl)
    for name, sid, t0, t1 in tl:
        if lo <= t0 - t_base < hi:
            print(f"  s{sid} {t0 - t_base:9.3f} -> {t1 - t_base:9.3f}  ({t1 - t0:6.3f})  {name}")


if __name__ == "__main__":
    main()
