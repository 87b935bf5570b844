import sys, random, cProfile, pstats, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from image_processor_pipeline_b200 import chain, synth, _native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
frames_h, bgs_h = bench.make_inputs(B, seed=1)
dev = torch.device('cuda:0')
f = torch.from_numpy(frames_h).to(dev); b = torch.from_numpy(bgs_h).to(dev)
params = chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **bench.CHAIN_KW)
for s in range(3):
    random.seed(s); chain.run_chain(f, b, params)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(5):
    random.seed(10 + s); chain.run_chain(f, b, params)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for s in range(5):
    random.seed(10 + s); chain.run_chain(f, b, params)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
