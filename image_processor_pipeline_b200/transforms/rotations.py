"""process_rotations — drop-in for reference transforms/rotations.py:6-133."""
from __future__ import annotations

import random
from pathlib import Path
from typing import Any, List, Optional

import numpy as np
from PIL import Image, UnidentifiedImageError

from .. import _io, ops


def process_rotations(input_path: Path, output_dirs: List[Path], num_rotations: int = 10,
                      include_original: bool = True, angle_min: float = 1.0, angle_max: float = 359.0,
                      output_format: str = "png", output_prefix: str = "r", original_key: str = "r000",
                      rotation_key_format: str = "{prefix}{index:03d}", resample: str = "nearest",
                      **options: Any) -> Optional[List[Path]]:
    """RGBA conversion, optional ``<stem>_r000`` original, then ``num_rotations`` times:
    one ``random.uniform`` angle, rotate(expand) + getbbox + crop on the GPU (one launch
    writes the canvas and reduces the alpha bbox), save ``<stem>_rNNN``.  Never raises:
    prints and returns None / the list, like the reference.  ``resample`` ('nearest' — the
    reference's mode — or 'bilinear') is an extra option."""
    if not output_dirs:
        print(f"Erreur [{input_path.name} - Rotation]: Aucun dossier de sortie ('output_paths') fourni.")
        return None
    target_dir = output_dirs[0]
    try:
        img = Image.open(input_path).convert("RGBA")
    except FileNotFoundError:
        print(f"Erreur [{input_path.name} - Rotation]: Fichier non trouvé.")
        return None
    except UnidentifiedImageError:
        print(f"Erreur [{input_path.name} - Rotation]: Impossible d'identifier ou d'ouvrir l'image (format invalide?).")
        return None
    except Exception as e:
        print(f"Erreur [{input_path.name} - Rotation]: Échec lors de la lecture du fichier: {e}")
        return None

    saved: List[Path] = []
    stem = input_path.stem
    suffix = ".jpg" if output_format.lower() == "jpeg" else f".{output_format.lower()}"
    if include_original:
        path = target_dir / f"{stem}_{original_key}{suffix}"
        try:
            img.save(path, format=output_format)
            saved.append(path)
        except Exception as e_save:
            print(f"Erreur [{input_path.name} - Rotation]: Échec sauvegarde de l'original '{path.name}': {e_save}")

    dev_img = None
    for i in range(num_rotations):
        angle = random.uniform(angle_min, angle_max)
        try:
            if dev_img is None:
                dev_img = _io.to_device(np.asarray(img))
            rotated, bbox = ops.rotate_and_fit(dev_img, angle, resample)
            if bbox is None:
                print(f"Avertissement [{input_path.name} - Rotation]: Impossible d'obtenir BBox après rotation {i+1}. "
                      "Utilisation de l'image non recadrée.")
            key = rotation_key_format.format(prefix=output_prefix, index=i + 1)
            path = target_dir / f"{stem}_{key}{suffix}"
            Image.fromarray(_io.to_host(rotated), "RGBA").save(path, format=output_format)
            saved.append(path)
        except Exception as e_rot:
            print(f"Erreur [{input_path.name} - Rotation]: Échec lors de la génération/sauvegarde de la rotation {i+1} "
                  f"(angle {angle:.1f}°): {e_rot}")
    if not saved:
        print(f"Avertissement [{input_path.name} - Rotation]: Aucune image (originale ou rotation) n'a pu être sauvegardée.")
        return None
    return saved


process_rotations.uses_gpu = True
