"""keep_largest_component — drop-in for reference transforms/pixels_isolés.py:8-81."""
from __future__ import annotations

from pathlib import Path
from typing import List

import cv2
import numpy as np

from .. import _io, ops
from ..utils.utils import _validate_dirs


def keep_largest_component(file: Path, output_dirs: List[Path], min_component_size: int = 500):
    """BGRA PNG -> alpha of everything outside the largest 8-connected component of
    alpha > 1 set to 0 -> tight crop to alpha != 0 -> same file name.
    Signature has no **options, like the reference.  ``min_component_size`` is accepted
    and, as in the reference (whose loop over it is a no-op), has no effect.
    Raises ValueError (not .png), FileNotFoundError (unreadable), AttributeError (no alpha;
    also for 2-D images, where the reference dies with IndexError).  Returns the Path, or
    None when the result is empty / cannot be written."""
    output_dir = _validate_dirs(output_dirs, nb_dirs=1)
    if file.suffix.lower() != ".png":
        raise ValueError(f"Le fichier {file.name} n'est pas un PNG.")
    image = cv2.imread(str(file), cv2.IMREAD_UNCHANGED)
    if image is None:
        raise FileNotFoundError(f"Impossible de charger l'image {file.name}.")
    if image.ndim != 3 or image.shape[2] != 4:
        raise AttributeError(f"L'image {file.name} ne contient pas de canal alpha, elle sera ignorée.")
    if image.dtype != np.uint8:
        raise NotImplementedError("keep_largest_component: 8-bit BGRA only on the GPU path")
    cleaned, info = ops.keep_largest_component(_io.to_device(image), inplace=True)
    x, y, w, h = (int(v) for v in info[0, :4])
    output_path = output_dir / file.name
    try:
        cropped = _io.to_host(ops.flip_crop(cleaned, x, y, w, h)) if w > 0 and h > 0 else image[0:0, 0:0]
        cv2.imwrite(str(output_path), cropped)      # raises on an empty image, like the reference
        return output_path
    except Exception as e_save:
        print(f"Erreur [{file.name}]: Échec de sauvegarde pour {output_path.name}: {e_save}")
        return None


def _crop_fit(img: np.ndarray) -> np.ndarray:
    """Tight crop of a host BGRA array to alpha != 0 using the device bbox reduction
    (reference pixels_isolés.py:74-81)."""
    bbox = ops.nonzero_bbox(_io.to_device(img), alpha_only=True)
    if bbox is None:
        return img[0:0, 0:0]
    x0, y0, x1, y1 = bbox
    return img[y0:y1, x0:x1]


keep_largest_component.uses_gpu = True
