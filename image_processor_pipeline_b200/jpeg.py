"""On-GPU JPEG codec either side of the chain: ctypes binding of libippjpeg (include/ippjpeg.h),
which drives NVIDIA's nvJPEG.  Replaces the host codec calls the reference makes for `.jpg` files
(video.py:44 writes them, recadrages.py:29 / symmetry.py:109 read them): the compressed bytes cross PCIe,
the pixels are born in HBM.  No CPU fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np
import torch

_LIB = None


def lib_path() -> Path:
    return Path(__file__).resolve().parent / "libippjpeg.so"


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.is_file():
        raise ImportError(f"{p} not found: build it with `python -m image_processor_pipeline_b200.build` (needs libnvjpeg). There is no CPU fallback.")
    lib = C.CDLL(str(p))
    lib.ippjpeg_last_error_string.restype = C.c_char_p
    lib.ippjpeg_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.ippjpeg_destroy.argtypes = [C.c_void_p]
    lib.ippjpeg_backend.argtypes = [C.c_void_p]
    lib.ippjpeg_info.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ippjpeg_decode_batch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.ippjpeg_encode_bound.restype = C.c_size_t
    lib.ippjpeg_encode_bound.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ippjpeg_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_size_t), C.c_void_p]
    _LIB = lib
    return lib


class JpegError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise JpegError(f"{what}: code {rc}: {load().ippjpeg_last_error_string().decode('utf-8', 'replace')}")


class JpegCodec:
    """One nvJPEG instance on the current device.  ``backend``: 0 = default (GPU-assisted Huffman for batches),
    1 = the hardware JPEG engine when available (``self.backend`` tells what was obtained)."""

    def __init__(self, device=None, backend: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("JpegCodec needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(load().ippjpeg_create(C.byref(self._h), int(backend)), "ippjpeg_create")
        self.backend = int(load().ippjpeg_backend(self._h))
        self._keep = []

    def close(self):
        if self._h:
            load().ippjpeg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, data) -> tuple:
        buf = np.frombuffer(data, dtype=np.uint8)
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        _check(load().ippjpeg_info(self._h, buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(c)), "ippjpeg_info")
        return w.value, h.value, c.value

    def decode_batch(self, datas: Sequence, w: int, h: int, bgr: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """n JPEG bitstreams (bytes / uint8 arrays, all w x h, 3 components) -> uint8 [n, h, w, 3] on the device (BGR like cv2.imread
        when ``bgr``), queued on the current stream.  The bitstreams are kept alive until the next call."""
        n = len(datas)
        if out is None:
            out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        elif tuple(out.shape) != (n, h, w, 3) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
            raise ValueError("out must be a contiguous CUDA uint8 tensor [n, h, w, 3]")
        bufs = [np.frombuffer(d, dtype=np.uint8) for d in datas]
        ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * max(n, 1))(*[b.size for b in bufs])
        self._keep = [bufs, ptrs, sizes]
        with torch.cuda.device(self.device):
            _check(load().ippjpeg_decode_batch(self._h, ptrs, sizes, n, w, h, int(bool(bgr)), out.data_ptr(), w * 3, h * w * 3,
                                               torch.cuda.current_stream().cuda_stream), "ippjpeg_decode_batch")
        return out

    def encode(self, img: torch.Tensor, quality: int = 95, bgr: bool = True) -> bytes:
        """uint8 [h, w, 3] device image -> baseline JPEG bytes (4:2:0 like cv2.imwrite's default)."""
        if img.dtype != torch.uint8 or not img.is_cuda or img.dim() != 3 or img.shape[2] != 3:
            raise ValueError("encode expects a CUDA uint8 [h, w, 3] image")
        img = img.contiguous()
        h, w, _ = img.shape
        cap = int(load().ippjpeg_encode_bound(self._h, w, h))
        buf = np.empty(cap, dtype=np.uint8)
        size = C.c_size_t(0)
        with torch.cuda.device(self.device):
            _check(load().ippjpeg_encode(self._h, img.data_ptr(), w * 3, w, h, int(bool(bgr)), int(quality), buf.ctypes.data, cap, C.byref(size),
                                         torch.cuda.current_stream().cuda_stream), "ippjpeg_encode")
        return buf[:size.value].tobytes()
