"""frame_extraction — drop-in for reference transforms/video.py:6-47, plus a device frame
source for the fused chain (frames go straight to HBM instead of JPEG files)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Iterator, List, Optional

import cv2
import numpy as np
import torch

from .. import _io


def frame_extraction(video_path: Path, output_dirs: List[Path], file_basename: str, **options: Any) -> Optional[Path]:
    """Every frame of the video -> ``<out>/<video stem>/0-raw/<basename>-frame_%04d.jpg``
    (1-based).  ValueError without a basename or for a non-video extension (checked after
    opening, like the reference), RuntimeError when the video cannot be opened.  Returns
    the frame DIRECTORY.  Decoding is the codec's job (NVDEC is a next row)."""
    if not file_basename:
        raise ValueError("Aucun nom de fichier de base fournit pour les nom des frames.")
    output_dir = output_dirs[0] / video_path.stem / "0-raw"
    output_dir.mkdir(parents=True, exist_ok=True)
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise RuntimeError("Erreur : Impossible d'ouvrir la vidéo")
    if video_path.suffix[1:].lower() not in _io.VID_FORMATS:
        raise ValueError(f"Fichier vidéo {video_path.suffix} non pris en charge. Format autorisés : {_io.VID_FORMATS}")
    frame_count = 1
    while True:
        success, frame = cap.read()
        if not success:
            break
        cv2.imwrite(str(output_dir / f"{file_basename}-frame_{frame_count:04d}.jpg"), frame)
        frame_count += 1
    cap.release()
    return output_dir


def frame_batches(video_path: Path, batch: int = 64, device: Optional[torch.device] = None) -> Iterator[torch.Tensor]:
    """Decoded frames as uint8 [B, H, W, 3] BGR CUDA tensors, staged through one pinned
    buffer — the input ``chain.run_chain`` consumes (no JPEG round trip)."""
    dev = device or _io.device()
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise RuntimeError("Erreur : Impossible d'ouvrir la vidéo")
    pinned = None
    n = 0
    try:
        while True:
            ok, frame = cap.read()
            if ok:
                if pinned is None:
                    pinned = torch.empty((batch,) + frame.shape, dtype=torch.uint8).pin_memory()
                pinned[n].copy_(torch.from_numpy(np.ascontiguousarray(frame)))
                n += 1
            if n == batch or (not ok and n > 0):
                yield pinned[:n].to(dev, non_blocking=True).clone()
                torch.cuda.current_stream(dev).synchronize()
                n = 0
            if not ok:
                break
    finally:
        cap.release()
