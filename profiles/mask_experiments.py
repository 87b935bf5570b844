"""Where does the stage-A mask kernel's time go?  Same frames, different filter lists / contents."""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from image_processor_pipeline_b200 import _native, chain, hostmath, synth
lib = _native.load()
dev = torch.device('cuda:0')
B = 64
frames_h, _ = bench.make_inputs(B, seed=1)
variants = {
    "synthetic frames, reference filters (3 hue filters)": (frames_h, synth.REFERENCE_FILTERS),
    "synthetic frames, filter 0 only (no hue filter)": (frames_h, synth.REFERENCE_FILTERS[:1]),
    "all-dark frames, reference filters": (np.full_like(frames_h, 40), synth.REFERENCE_FILTERS),
    "uniform noise frames, reference filters": (np.random.default_rng(0).integers(0, 256, frames_h.shape, dtype=np.uint8), synth.REFERENCE_FILTERS),
    "uniform noise frames, filter 0 only": (np.random.default_rng(0).integers(0, 256, frames_h.shape, dtype=np.uint8), synth.REFERENCE_FILTERS[:1]),
}
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, (fr, filt) in variants.items():
    if only and only not in name:
        continue
    f = torch.from_numpy(fr).to(dev)
    geom = chain._geom_struct(1920, 1080, (0, 0, 1920, 1080), 1920, 1080, hostmath.pack_hsv_ranges(filt))
    ws_bytes = lib.ipp_chain_mask_workspace_bytes(C.byref(geom), B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    info = torch.empty((B, 8), dtype=torch.int32, device=dev)
    for it in range(3):
        if it == 2:
            _native.timing_enable(True)
        _native.check(lib.ipp_chain_stage_a(f.data_ptr(), 1920 * 3, 1080 * 1920 * 3, C.byref(geom), info.data_ptr(), ws.data_ptr(), ws_bytes, B,
                                            torch.cuda.current_stream().cuda_stream), "a")
    torch.cuda.synchronize()
    rep = _native.timing_report(); _native.timing_enable(False)
    kname = [k for k in rep if k.startswith("chain_mask_")][0]
    ms = rep[kname][1]
    print(f"{name:60s} {kname} {ms:.4f} ms  -> {B*6.2208e6/ms/1e6:.0f} GB/s   ccl: " + " ".join(f"{k[4:-7]}={v[1]:.3f}" for k, v in rep.items() if k.startswith("ccl_")))
