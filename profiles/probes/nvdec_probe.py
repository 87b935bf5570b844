import ctypes, glob, os, subprocess, sys
import numpy as np, cv2
print("cv2", cv2.__version__)
bi = cv2.getBuildInformation()
for line in bi.splitlines():
    if any(k in line for k in ("FFMPEG", "avcodec", "avformat", "GStreamer")):
        print(line.strip())
print([p for p in glob.glob("/usr/lib/x86_64-linux-gnu/libnvcuvid*") + glob.glob("/usr/lib64/libnvcuvid*") + glob.glob("/usr/local/cuda/lib64/*cuvid*")])
print(subprocess.run(r"ldconfig -p | grep -i 'cuvid\|nvidia-encode\|nvjpeg' | head", shell=True, capture_output=True, text=True).stdout)
for name in ("libnvcuvid.so.1", "libnvcuvid.so"):
    try:
        lib = ctypes.CDLL(name); print("loaded", name, [hasattr(lib, f) for f in ("cuvidCreateVideoParser", "cuvidCreateDecoder", "cuvidGetDecoderCaps", "cuvidMapVideoFrame64")]); break
    except OSError as e:
        print("no", name, e)
# can cv2 write videos here, with which codecs, and read raw packets back?
os.makedirs("/tmp/v", exist_ok=True)
for fourcc, ext in (("mp4v", "mp4"), ("avc1", "mp4"), ("H264", "mp4"), ("hev1", "mp4"), ("MJPG", "avi"), ("VP90", "webm")):
    path = f"/tmp/v/t_{fourcc}.{ext}"
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), 25.0, (320, 240))
    ok = w.isOpened()
    if ok:
        for i in range(5):
            w.write(np.full((240, 320, 3), 40 * i, np.uint8))
    w.release()
    size = os.path.getsize(path) if os.path.exists(path) else -1
    raw = None
    if ok and size > 0:
        cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
        print("  codec fourcc read back:", int(cap.get(cv2.CAP_PROP_FOURCC)).to_bytes(4, "little"), "raw-mode set:", cap.set(cv2.CAP_PROP_FORMAT, -1))
        r, pkt = cap.read()
        raw = (r, None if pkt is None else (pkt.shape, pkt.dtype, bytes(pkt.reshape(-1)[:8]).hex()))
        extra = cap.get(cv2.CAP_PROP_CODEC_EXTRADATA_INDEX) if hasattr(cv2, "CAP_PROP_CODEC_EXTRADATA_INDEX") else None
        print("  extradata index", extra)
        cap.release()
    print(fourcc, "writer", ok, "bytes", size, "raw packet", raw)
print(subprocess.run(r"nvidia-smi -q | grep -i -A3 'decoder\|FBC\|Encoder' | head -30", shell=True, capture_output=True, text=True).stdout)
