"""crop_from_border / fit_crop — drop-ins for reference transforms/recadrages.py:7-82."""
from __future__ import annotations

from pathlib import Path
from typing import Any, List, Optional, Tuple

import cv2
import numpy as np
from PIL import Image

from .. import _io, hostmath, ops

_compute_crop = hostmath.compute_crop


def crop_from_border(file: Path, output_dirs: List[Path], crop_margins: Tuple[float, float, float, float] = (0, 0, 0, 0),
                     **options: Any) -> Optional[Path]:
    """JPEG only; margins (top, bottom, left, right) as a fraction (<1) or pixels;
    ValueError for negative / oversized margins, FileNotFoundError when unreadable
    (reference recadrages.py:23-46).  The crop itself is a device address remap."""
    output_dir = output_dirs[0]
    if file.suffix.lower() not in (".jpg", ".jpeg"):
        raise ValueError(f"Le Fichier {file.name} n'est pas du type JPG.")
    image = cv2.imread(str(file), cv2.IMREAD_UNCHANGED)
    if image is None:
        raise FileNotFoundError(f"Impossible de charger l'image {file.name}.")
    height, width = image.shape[:2]
    x0, y0, w, h = hostmath.border_crop_rect(width, height, crop_margins, file.name)
    cropped = _io.to_host(ops.flip_crop(_io.to_device(image), x0, y0, w, h), like=image)
    output_path = output_dir / file.name
    try:
        if cv2.imwrite(str(output_path), cropped):
            return output_path
        print(f"Avertissement [{file.name}]: Échec de sauvegarde (imwrite a retourné False) pour {output_path.name}")
        return None
    except Exception as e_save:
        print(f"Erreur [{file.name}]: Échec de sauvegarde pour {output_path.name}: {e_save}")
        return None


def fit_crop(image_path: Path, output_dirs: List[Path], **options: Any) -> Optional[Path]:
    """Remove the transparent (or all-zero) border: PIL getbbox + crop, copy when empty
    (reference recadrages.py:63-82).  The bbox is a device reduction."""
    output_dir = output_dirs[0]
    image = Image.open(image_path)
    if image.mode not in ("L", "LA", "RGB", "RGBA"):
        raise NotImplementedError(f"fit_crop: PIL mode {image.mode} is not supported on the GPU path")
    arr = np.asarray(image)
    dev = _io.to_device(arr)
    bbox = ops.nonzero_bbox(dev)
    if not bbox:
        new_image = image.copy()
    else:
        x0, y0, x1, y1 = bbox
        out = _io.to_host(ops.flip_crop(dev, x0, y0, x1 - x0, y1 - y0), like=arr)
        new_image = Image.fromarray(out, image.mode)
    output_path = output_dir / image_path.name
    new_image.save(output_path)
    return output_path


crop_from_border.uses_gpu = True
fit_crop.uses_gpu = True
