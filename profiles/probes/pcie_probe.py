"""PCIe probe for the host-buffer entry point: contiguous H2D, 2-D window H2D (cudaMemcpy2DAsync per image),
D2H, and the three together (what run_chain_host overlaps).  GB/s, pinned memory, one GPU."""
import ctypes as C, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from image_processor_pipeline_b200 import _native, chain
lib = _native.load()
dev = torch.device('cuda:0')
N, H, W, CS = 64, 1080, 1920, 1080
frames = torch.empty((N, H, W, 3), dtype=torch.uint8).pin_memory()
bgs = torch.empty((N, H, W, 3), dtype=torch.uint8).pin_memory()
outs = torch.empty((N, CS, CS, 3), dtype=torch.uint8).pin_memory()
d_frames = torch.empty((N, H, W, 3), dtype=torch.uint8, device=dev)
d_win = torch.empty((N, CS, CS, 3), dtype=torch.uint8, device=dev)
d_out = torch.empty((N, CS, CS, 3), dtype=torch.uint8, device=dev)
items = np.zeros(N, dtype=chain.ITEM_DTYPE)
items["bg"] = np.arange(N); items["sq_size"] = CS; items["sq_x0"] = np.random.default_rng(0).integers(0, W - CS, N); items["sq_y0"] = 0
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_frames.copy_(frames, non_blocking=True)


def win():
    _native.check(lib.ipp_chain_upload_windows(bgs.data_ptr(), W * 3, H * W * 3, W, H, N, items.ctypes.data, N, CS, d_win.data_ptr(), s2.cuda_stream), "win")


def d2h():
    with torch.cuda.stream(s3):
        outs.copy_(d_out, non_blocking=True)


def timed(fns, nbytes, reps=5):
    for f in fns: f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for f in fns: f()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return [b / dt / 1e9 for b in nbytes], dt * 1e3


fb, wb, ob = frames.numel(), N * CS * CS * 3, outs.numel()
print("H2D contiguous        GB/s", timed([h2d], [fb]))
print("H2D 2-D windows       GB/s", timed([win], [wb]))
print("D2H contiguous        GB/s", timed([d2h], [ob]))
print("H2D + windows (2 str) GB/s", timed([h2d, win], [fb + wb]))
print("H2D + D2H             GB/s", timed([h2d, d2h], [fb, ob]))
print("all three             GB/s (h2d total, d2h)", timed([h2d, win, d2h], [fb + wb, ob]))
