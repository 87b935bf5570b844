"""process_images_with_color_masks — drop-in for reference transforms/filtres_liste.py:8-149."""
from __future__ import annotations

from pathlib import Path
from typing import Any, List, Optional, Tuple

import cv2

from .. import _io, hostmath, ops
from ..utils.utils import _validate_dirs

_rescale_filter = hostmath.rescale_filter


def process_images_with_color_masks(image_path: Path, output_dirs: List[Path],
                                    color_ranges_to_exclude_hsv: List[Tuple[int, int, int, int, int, int]],
                                    zones: Optional[List[Optional[Tuple[int, int, int, int]]]] = None,
                                    use_gimp_scale: bool = False, output_prefix: str = "",
                                    **options: Any) -> Optional[Path]:
    """BGR image -> BGRA PNG whose alpha is 0 wherever the pixel falls in ANY of the HSV
    ranges (inside that range's zone), 255 elsewhere; name ``[prefix_]<stem>.png``.
    ValueError for an empty filter list / mismatching zones / malformed filters, IOError
    when the image cannot be read (reference filtres_liste.py:74-86).  BGR->HSV, the range
    tests and the BGRA merge are one GPU kernel, bit-exact with OpenCV's integer HSV."""
    output_dir = _validate_dirs(output_dirs, nb_dirs=1)
    if not color_ranges_to_exclude_hsv:
        raise ValueError(f"Erreur [{image_path.name} - ColorMask] : `color_ranges_to_exclude_hsv` est requis pour traiter les données")
    if zones and len(zones) != len(color_ranges_to_exclude_hsv):
        raise ValueError(f"Les zones d'application des filtres colorimétriques ({len(zones)}) ne correspondent pas aux "
                         f"filtres ({len(color_ranges_to_exclude_hsv)}). Les 2 paramètres doivent être de même longueur !.")
    image = cv2.imread(str(image_path))
    if image is None:
        raise IOError("Impossible de charger l'image.")
    result = ops.color_masks_to_alpha(_io.to_device(image), color_ranges_to_exclude_hsv, zones, use_gimp_scale)
    output_filename = f"{output_prefix}{'_' if output_prefix else ''}{image_path.stem}.png"
    output_path = output_dir / output_filename
    try:
        if cv2.imwrite(str(output_path), _io.to_host(result)):
            return output_path
        raise RuntimeError(f"Échec de sauvegarde (imwrite a retourné False) pour {output_filename}")
    except Exception as e_save:
        print(f"Erreur lors de la sauvegarde de {output_path}: {e_save}")
        return None


process_images_with_color_masks.uses_gpu = True
