"""Host-side phase times of the end-to-end entry (run_chain_host, pinned host buffers): where does the host wait?
python profiles/prof_host_e2e.py [frames] [chunk]"""
import collections, random, sys, time
sys.path.insert(0, '.')
import torch
import bench
from image_processor_pipeline_b200 import chain, synth
E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device('cuda:0')
f, b = bench.make_inputs_device(E, 1, 0, dev)
fp = torch.empty(f.shape, dtype=torch.uint8).pin_memory(); fp.copy_(f)
bp = torch.empty(b.shape, dtype=torch.uint8).pin_memory(); bp.copy_(b)
out = torch.empty((E, 1080, 1080, 3), dtype=torch.uint8).pin_memory()
del f, b
params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **bench.CHAIN_KW)
for s in range(2):
    random.seed(s); chain.run_chain_host(fp, bp, params, dev, out, chunk=chunk)
chain.TRACE = []
t0 = time.perf_counter()
N = 3
for s in range(N):
    random.seed(10 + s)
    chain._mark("call")
    r = chain.run_chain_host(fp, bp, params, dev, out, chunk=chunk)
    chain._mark("returned")
dt = (time.perf_counter() - t0) / N
print(f"{E} frames, chunk {chunk}: {dt * 1e3:.1f} ms per call -> {E / dt:.0f} img/s; h2d {r.h2d_bytes / 1e9:.2f} GB -> {r.h2d_bytes / dt / 1e9:.1f} GB/s")
acc = collections.OrderedDict()
prev = None
for name, ts in chain.TRACE:
    if prev is not None:
        acc.setdefault(name, []).append(ts - prev)
    prev = ts
for name, v in acc.items():
    print(f"  -> {name:24s} {1e3 * sum(v) / len(v):7.3f} ms avg  x{len(v) // N:4d} per call  = {1e3 * sum(v) / N:7.2f} ms per call")
