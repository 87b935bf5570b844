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
    arr = np.asarray(image)
    if image.mode in ("L", "LA", "RGB", "RGBA"):
        dev = _io.to_device(arr)
        bbox = ops.nonzero_bbox(dev)
        if not bbox:
            new_image = image.copy()
        else:
            x0, y0, x1, y1 = bbox
            out = _io.to_host(ops.flip_crop(dev, x0, y0, x1 - x0, y1 - y0), like=arr)
            new_image = Image.fromarray(out, image.mode)
    else:
        # any other PIL mode (P, PA, 1, CMYK, YCbCr, I;16, I, F, La, RGBa ...): getbbox tests the alpha band for modes that have
        # one and "any byte of the pixel non-zero" otherwise (libImaging GetBBox.c works on the raw pixel bytes) — the same
        # device reduction on the byte view; the crop itself stays with PIL so that palette / info travel like in the reference
        raw = np.ascontiguousarray(arr.astype(np.uint8) if arr.dtype == bool else arr)
        byts = raw.view(np.uint8).reshape(raw.shape[0], raw.shape[1], -1)
        if image.mode.startswith("I;16"):
            # libImaging keeps I;16 rows in its 8-bit plane and GetBBox.c scans `xsize` BYTES of each row there: Pillow's bbox of a
            # 16-bit image is the bbox of the first `width` bytes of every row, one byte = one column.  Reproduced as is.
            byts = np.ascontiguousarray(byts.reshape(raw.shape[0], -1)[:, :raw.shape[1], None])
        alpha_only = image.mode in ("PA", "La", "RGBa")
        bbox = ops.nonzero_bbox(_io.to_device(byts), alpha_only=alpha_only)
        new_image = image.crop(bbox) if bbox else image.copy()
    output_path = output_dir / image_path.name
    new_image.save(output_path)
    return output_path


crop_from_border.uses_gpu = True
fit_crop.uses_gpu = True
