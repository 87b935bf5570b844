"""NVDEC frame source in front of the chain: ctypes binding of libippvideo (include/ippvideo.h), which drives
the GPU's video decoder through the driver's libnvcuvid.  Replaces the host decode of reference transforms/video.py:23-45
(cv2.VideoCapture.read per frame, then a .jpg round trip): the compressed packets cross PCIe, the frames are born in HBM as
the [B, H, W, 3] BGR batches ``chain.run_chain`` consumes.  Demuxing stays on the host (cv2.VideoCapture in raw-packet mode).
No CPU fallback: a missing library / an unsupported codec raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Iterator, Optional

import numpy as np
import torch

_LIB = None

# cudaVideoCodec values (include/ippvideo.h) by the FOURCC cv2 reports for the stream
CODECS = {"vp90": 10, "vp09": 10, "vp80": 9, "fmp4": 2, "mp4v": 2, "xvid": 2, "divx": 2, "avc1": 4, "h264": 4, "x264": 4,
          "hev1": 8, "hvc1": 8, "hevc": 8, "av01": 11, "mpg1": 0, "mpg2": 1, "mpeg": 1}


def lib_path() -> Path:
    return Path(__file__).resolve().parent / "libippvideo.so"


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.is_file():
        raise ImportError(f"{p} not found: build it with `python -m image_processor_pipeline_b200.build`. There is no CPU fallback.")
    lib = C.CDLL(str(p))
    lib.ippvideo_last_error_string.restype = C.c_char_p
    lib.ippvideo_supported.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ippvideo_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.ippvideo_destroy.argtypes = [C.c_void_p]
    lib.ippvideo_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ippvideo_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    _LIB = lib
    return lib


class VideoError(RuntimeError):
    pass


def _err() -> str:
    return load().ippvideo_last_error_string().decode("utf-8", "replace")


def supported(codec: int) -> bool:
    """True when libnvcuvid is present and this GPU's NVDEC decodes ``codec`` (8-bit 4:2:0)."""
    if not torch.cuda.is_available():
        return False
    torch.cuda.init()
    torch.zeros(1, device="cuda")                      # the primary context is current on this thread
    return bool(load().ippvideo_supported(int(codec), None, None))


class VideoDecoder:
    """One NVDEC parser + decoder for one elementary stream on the current device."""

    def __init__(self, codec: int, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("VideoDecoder needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)
            rc = load().ippvideo_create(C.byref(self._h), int(codec))
        if rc != 0:
            raise VideoError(f"ippvideo_create: code {rc}: {_err()}")

    def decode(self, packet, out: torch.Tensor, end_of_stream: bool = False) -> int:
        """Feed one packet (bytes / uint8 array; empty + ``end_of_stream`` flushes); frames that become displayable are written to
        ``out`` [capacity, H, W, 3] (uint8 CUDA, BGR).  Returns their number."""
        if out.dtype != torch.uint8 or not out.is_cuda or out.dim() != 4 or out.shape[3] != 3 or out.stride(3) != 1 or out.stride(2) != 3:
            raise ValueError("out must be a uint8 CUDA tensor [capacity, H, W, 3] with dense pixels")
        buf = np.ascontiguousarray(np.frombuffer(packet, dtype=np.uint8) if isinstance(packet, (bytes, bytearray, memoryview)) else
                                   np.asarray(packet, dtype=np.uint8).reshape(-1))
        n, w, h = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(self.device):
            rc = load().ippvideo_decode(self._h, buf.ctypes.data if buf.size else None, buf.size, int(end_of_stream), out.data_ptr(),
                                        out.stride(1), out.stride(0), out.shape[0], C.byref(n), C.byref(w), C.byref(h),
                                        torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise VideoError(f"ippvideo_decode: code {rc}: {_err()}")
        if n.value and (w.value != out.shape[2] or h.value != out.shape[1]):
            raise VideoError(f"the stream is {w.value} x {h.value}, the output buffer {out.shape[2]} x {out.shape[1]}")
        return n.value

    def close(self):
        if self._h:
            load().ippvideo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _open_raw(video_path):
    """cv2.VideoCapture in raw-packet mode (the FFmpeg backend hands out the demuxed access units; H.264 / HEVC already
    in Annex B form).  Returns (capture, codec, width, height, extradata or None)."""
    import cv2
    cap = cv2.VideoCapture(str(video_path), cv2.CAP_FFMPEG)
    if not cap.isOpened():
        raise RuntimeError("Erreur : Impossible d'ouvrir la vidéo")
    fourcc = int(cap.get(cv2.CAP_PROP_FOURCC)).to_bytes(4, "little").decode("ascii", "replace").lower()
    w, h = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    if fourcc not in CODECS:
        cap.release()
        raise VideoError(f"codec {fourcc!r} is not one NVDEC decodes")
    if not cap.set(cv2.CAP_PROP_FORMAT, -1):
        cap.release()
        raise VideoError("this cv2 build cannot hand out raw packets (CAP_PROP_FORMAT = -1)")
    extra = None
    idx = int(cap.get(cv2.CAP_PROP_CODEC_EXTRADATA_INDEX)) if hasattr(cv2, "CAP_PROP_CODEC_EXTRADATA_INDEX") else -1
    if idx >= 0:
        ok, blob = cap.retrieve(flag=idx)
        if ok and blob is not None and blob.size:
            extra = np.ascontiguousarray(blob).reshape(-1).copy()
    return cap, CODECS[fourcc], w, h, extra


def frame_batches(video_path, batch: int = 64, device: Optional[torch.device] = None) -> Iterator[torch.Tensor]:
    """Decoded frames of a video as uint8 [B, H, W, 3] BGR CUDA tensors (B <= batch), decoded by NVDEC: what
    ``transforms.video.frame_batches`` yields, without the frames ever existing on the host."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    cap, codec, w, h, extra = _open_raw(video_path)
    dec = None
    try:
        dec = VideoDecoder(codec, dev)
        slack = 20                                          # frames one packet may release at once (reordering depth, flush)
        buf = torch.empty((batch + slack, h, w, 3), dtype=torch.uint8, device=dev)
        n = 0
        if extra is not None and codec in (2,):             # MPEG-4 part 2: the VOL header lives in the container
            n += dec.decode(extra, buf[n:])
        while True:
            ok, pkt = cap.read()
            if not ok or pkt is None or pkt.size == 0:
                break
            n += dec.decode(pkt, buf[n:])
            while n >= batch:
                yield buf[:batch].clone()
                rest = n - batch
                if rest:
                    buf[:rest].copy_(buf[batch:n].clone())
                n = rest
        while True:                                         # flush: the decoder releases what it still holds
            got = dec.decode(b"", buf[n:], end_of_stream=True)
            n += got
            while n >= batch:
                yield buf[:batch].clone()
                rest = n - batch
                if rest:
                    buf[:rest].copy_(buf[batch:n].clone())
                n = rest
            if got == 0:
                break
        if n:
            yield buf[:n].clone()
    finally:
        cap.release()
        if dec is not None:
            dec.close()
