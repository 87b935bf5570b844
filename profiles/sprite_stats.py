"""Sizes the LANCZOS passes see in the bench workload (sprite, resized sprite, kernel support, super-blocks per tile)."""
import sys
sys.path.insert(0, '.')
import random
import numpy as np, torch
import bench
from image_processor_pipeline_b200 import chain

cap = []
orig = chain._layout
def spy(items):
    r = orig(items)
    cap.append(items.copy())
    return r
chain._layout = spy
B = 64
from image_processor_pipeline_b200 import synth
frames, bgs = bench.make_inputs_device(B, 1, 0, torch.device('cuda:0'))
random.seed(0)
chain.run_chain(frames, bgs, chain.ChainParams(color_ranges_to_exclude_hsv=synth.REFERENCE_FILTERS, **bench.CHAIN_KW))
it = np.concatenate(cap)
for n in ("spr_w", "spr_h", "new_w", "new_h", "h_ksize", "v_ksize"):
    v = it[n].astype(np.int64)
    print(f"{n:8s} min {v.min():5d} mean {v.mean():8.1f} max {v.max():5d}")
print("nkb h", np.bincount(it["nkb"] & 0xffff), " v", np.bincount(it["nkb"] >> 16))
s = it["new_w"] / it["spr_w"]
print("scale  min %.3f mean %.3f max %.3f" % (s.min(), s.mean(), s.max()))
# columns of the 8-wide tiles actually covered by the band, in 32-blocks: span = ksize + 7 / scale (+15 for 16-alignment of xb)
span = it["h_ksize"] + 7 / s + 15
print("k32 blocks needed per tile (h): mean %.2f of %.2f issued" % (np.ceil(span / 32).mean(), 2 * (it["nkb"] & 0xffff).mean()))
