"""Box-level PCIe ceiling of the end-to-end entry point: every rank (one per GPU, torchrun) moves the traffic mix of
run_chain_host at the same time — frames H2D (contiguous) + crop windows H2D (2-D copies) + outputs D2H, three streams —
and the images/s the box allows is derived from the slowest rank.  1080p: 9.72 MB in + 3.50 MB out per image.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/probes/pcie_probe_multi.py
"""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import torch.distributed as dist
from image_processor_pipeline_b200 import _native, chain

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("gloo")
lib = _native.load()
N, H, W, CS = 48, 1080, 1920, 1080
frames = torch.empty((N, H, W, 3), dtype=torch.uint8).pin_memory()
bgs = torch.empty((N, H, W, 3), dtype=torch.uint8).pin_memory()
outs = torch.empty((N, CS, CS, 3), dtype=torch.uint8).pin_memory()
d_frames = torch.empty((N, H, W, 3), dtype=torch.uint8, device=dev)
d_win = torch.empty((N, CS, CS, 3), dtype=torch.uint8, device=dev)
d_out = torch.empty((N, CS, CS, 3), dtype=torch.uint8, device=dev)
items = np.zeros(N, dtype=chain.ITEM_DTYPE)
items["bg"] = np.arange(N); items["sq_size"] = CS; items["sq_x0"] = np.random.default_rng(rank).integers(0, W - CS, N)
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def mix():
    with torch.cuda.stream(s1):
        d_frames.copy_(frames, non_blocking=True)
    _native.check(lib.ipp_chain_upload_windows(bgs.data_ptr(), W * 3, H * W * 3, W, H, N, items.ctypes.data, N, CS, d_win.data_ptr(), s2.cuda_stream), "win")
    with torch.cuda.stream(s3):
        outs.copy_(d_out, non_blocking=True)


for _ in range(2):
    mix()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 8
t0 = time.perf_counter()
for _ in range(reps):
    mix()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
ips = N / dt
if world > 1:
    t = torch.tensor([ips], dtype=torch.float64)
    allv = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allv, t)
    per = [float(v) for v in allv]
else:
    per = [ips]
if rank == 0:
    h2d = frames.numel() + N * CS * CS * 3
    print(json.dumps({"n_gpus": world, "images_per_s_per_rank": [round(v, 1) for v in per], "box_images_per_s": round(world * min(per), 1),
                      "sum_images_per_s": round(sum(per), 1), "h2d_gbs_rank0": round(h2d / dt / 1e9, 2), "d2h_gbs_rank0": round(outs.numel() / dt / 1e9, 2),
                      "mix": "per image 6.22 MB frame + 3.50 MB crop window H2D, 3.50 MB result D2H, three streams per rank, all ranks at once"}))
if world > 1:
    dist.destroy_process_group()
