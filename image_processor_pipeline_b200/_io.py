"""Host <-> device plumbing of the file-based drop-in transforms.

Decode / encode stay on the host with the very library calls the reference uses
(cv2.imread / cv2.imwrite / PIL open / save): codecs are outside the hot path
 and keeping them identical makes the written files byte-identical
whenever the pixels are.  Everything between decode and encode runs on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

# extension sets the reference imports from ultralytics.data.utils (symmetry.py:5, video.py:4)
IMG_FORMATS = {"bmp", "dng", "jpeg", "jpg", "mpo", "png", "tif", "tiff", "webp", "pfm", "heic"}
VID_FORMATS = {"asf", "avi", "gif", "m4v", "mkv", "mov", "mp4", "mpeg", "mpg", "ts", "wmv", "webm"}


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("image_processor_pipeline_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(arr: np.ndarray) -> torch.Tensor:
    """uint8 HWC / HW array -> CUDA tensor (uint16 images travel as uint8 with doubled channel bytes)."""
    a = np.ascontiguousarray(arr)
    if a.dtype == np.uint16:
        a = a.view(np.uint8).reshape(a.shape[0], a.shape[1], -1)
    elif a.dtype != np.uint8:
        raise NotImplementedError(f"unsupported image dtype {a.dtype}")
    return torch.from_numpy(a).to(device(), non_blocking=False)


def to_host(t: torch.Tensor, like: np.ndarray | None = None) -> np.ndarray:
    out = t.cpu().numpy()
    if like is not None and like.dtype == np.uint16:
        out = out.view(np.uint16)
        if like.ndim == 2:
            out = out.reshape(out.shape[0], out.shape[1])
    elif like is not None and like.ndim == 2 and out.ndim == 3:
        out = out[..., 0]
    return out
