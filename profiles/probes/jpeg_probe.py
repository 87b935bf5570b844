"""nvJPEG decode throughput / fidelity probe: synthetic 1080p frames + noise backgrounds encoded by cv2 (libjpeg-turbo, the
codec the reference uses), decoded in batches on the GPU.  python profiles/probes/jpeg_probe.py"""
import sys, time
sys.path.insert(0, '.')
import cv2, numpy as np, torch
from image_processor_pipeline_b200 import jpeg, synth
H, W = 1080, 1920
imgs = [synth.frame(1, k, H, W) for k in range(4)] + [synth.background(1, k, H, W) for k in range(2)]
smooth = cv2.GaussianBlur(synth.background(1, 9, H, W), (0, 0), 6)          # a "natural image" stand-in (low-pass noise)
imgs.append(smooth)
for q in (95, 80):
    enc = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes() for im in imgs]
    print(f"quality {q}: bytes per image frames {np.mean([len(e) for e in enc[:4]]) / 1e6:.2f} MB, noise backgrounds {np.mean([len(e) for e in enc[4:6]]) / 1e6:.2f} MB, smooth {len(enc[6]) / 1e6:.2f} MB (raw 6.22 MB)")
    for backend in (2, 2 | 256, 256):
        codec = jpeg.JpegCodec(backend=backend)
        for name, sel in (("frames", enc[:4]), ("noise bg", enc[4:6]), ("smooth", enc[6:7])):
            batch = [sel[i % len(sel)] for i in range(64)]
            out = torch.empty((64, H, W, 3), dtype=torch.uint8, device="cuda")
            try:
                codec.decode_batch(batch, W, H, True, out)
            except Exception as e:
                print(f"  backend {backend} (got {codec.backend}) {name}: FAILED {e}")
                continue
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                codec.decode_batch(batch, W, H, True, out)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            ref = cv2.imdecode(np.frombuffer(batch[0], np.uint8), cv2.IMREAD_COLOR)
            d = np.abs(out[0].cpu().numpy().astype(np.int16) - ref.astype(np.int16))
            print(f"  backend {backend} (got {codec.backend}) {name:9s}: {64 / dt:8.0f} img/s   max |diff| vs cv2.imdecode {d.max()}, mean {d.mean():.4f}, >1: {(d > 1).mean():.5f}")
        codec.close()
codec = jpeg.JpegCodec()
x = torch.from_numpy(imgs[0]).cuda()
t0 = time.perf_counter()
for _ in range(10):
    b = codec.encode(x, 95)
dt = (time.perf_counter() - t0) / 10
back = cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR)
print(f"encode 1080p q95: {1 / dt:.0f} img/s, {len(b) / 1e6:.2f} MB, PSNR vs source {cv2.PSNR(back, imgs[0]):.1f} dB (cv2 encoder: {cv2.PSNR(cv2.imdecode(np.frombuffer(cv2.imencode('.jpg', imgs[0], [cv2.IMWRITE_JPEG_QUALITY, 95])[1], np.uint8), cv2.IMREAD_COLOR), imgs[0]):.1f} dB)")
