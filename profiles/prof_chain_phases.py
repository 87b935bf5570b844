"""Host-side phase timeline of the pipelined chain (ChainStream): where does the host wait, how long
does it plan?  python profiles/prof_chain_phases.py [batch]"""
import collections, random, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from image_processor_pipeline_b200 import chain, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
frames_h, bgs_h = bench.make_inputs(B, seed=1)
dev = torch.device('cuda:0')
f = torch.from_numpy(frames_h).to(dev); b = torch.from_numpy(bgs_h).to(dev)
params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **bench.CHAIN_KW)
pipe = chain.ChainStream(params)


def run(count, first):
    for _ in pipe.map(((f, b) for _ in range(count)), random, on_batch=lambda i: random.seed(first + i)):
        chain._mark("yield")


run(5, 0)
torch.cuda.synchronize()
chain.TRACE = []
t0 = time.perf_counter()
N = 8
run(N, 10)
torch.cuda.synchronize()
print(f"ms/step {(time.perf_counter() - t0) / N * 1e3:.3f}")
acc = collections.OrderedDict()
prev = None
for name, ts in chain.TRACE:
    if prev is not None:
        acc.setdefault(name, []).append(ts - prev)
    prev = ts
for name, v in acc.items():
    print(f"  -> {name:24s} {1e3 * sum(v) / len(v):7.3f} ms  (x{len(v)})")
