"""paste_overlay_onto_background — drop-in for reference transforms/overlays.py:24-187."""
from __future__ import annotations

import random
from pathlib import Path
from typing import Any, List, Optional, Tuple

import numpy as np
from PIL import Image, UnidentifiedImageError

from .. import _io, hostmath, ops
from ..utils import utils


def _convert_to_yolo_bbox(img_width: int, img_height: int, box: Tuple[int, int, int, int]):
    """reference overlays.py:13-22 (helper of the deprecated twin; kept for API parity)."""
    if img_width <= 0 or img_height <= 0:
        raise ValueError(f"Les dimensions de l'image ({img_width}x{img_height}) doivent être positives.")
    dw, dh = 1.0 / img_width, 1.0 / img_height
    return ((box[0] + box[2]) / 2.0 * dw, (box[1] + box[3]) / 2.0 * dh, (box[2] - box[0]) * dw, (box[3] - box[1]) * dh)


def paste_overlay_onto_background(overlay_path: Path, background_path: Path, output_dirs: List[Path],
                                  yolo_class_id: int = 0, scale_min: float = 0.15, scale_max: float = 0.30,
                                  **options: Any) -> Optional[List[Path]]:
    """Sprite -> RGBA, background -> RGB; size from a random ratio of the background
    diagonal (``random.uniform``), LANCZOS resize, random position (two ``random.randint``),
    alpha paste, YOLO label ``<cls> cx cy w h`` (6 decimals, no trailing newline).
    Outputs ``<sprite stem><bg suffix>`` and ``<sprite stem>.txt``.  IndexError from
    _validate_dirs; everything else prints and returns None, like the reference.
    Resize (two integer passes) and paste (DIV255 blend) run on the GPU, bit-exact."""
    image_target_dir, label_target_dir = utils._validate_dirs(output_dirs, nb_dirs=2)[:2]
    try:
        overlay = Image.open(overlay_path)
        if overlay.mode != "RGBA":
            overlay = overlay.convert("RGBA")
        background = Image.open(background_path).convert("RGB")
    except FileNotFoundError as fnf:
        print(f"Erreur [{overlay_path.name} + {background_path.name}]: Fichier non trouvé: {fnf}")
        return None
    except UnidentifiedImageError as uie:
        print(f"Erreur [{overlay_path.name} + {background_path.name}]: Impossible d'ouvrir l'image {uie}")
        return None
    except TypeError as te:
        print(f"Erreur [{overlay_path.name} + {background_path.name}]: Type d'image invalide : {te}")
        return None
    except Exception as e:
        print(f"Erreur [{overlay_path.name} + {background_path.name}]: Échec lecture fichiers: {e}")
        return None

    try:
        target_ratio = random.uniform(scale_min, scale_max)
        new_w, new_h = hostmath.overlay_size(overlay.width, overlay.height, background.width, background.height, target_ratio)
        dev = _io.device()
        resized = ops.resize_lanczos(_io.to_device(np.asarray(overlay)), (new_w, new_h))
        pos_x = random.randint(0, background.width - new_w)
        pos_y = random.randint(0, background.height - new_h)
        composite = ops.paste(_io.to_device(np.asarray(background)), resized, (pos_x, pos_y))
        box = np.array([pos_x, pos_y, pos_x + new_w, pos_y + new_h]).reshape(1, 4)
        cx, cy, w_norm, h_norm = ops.xyxy2xywhn(box, background.width, background.height, dev)[0]
        yolo_label_str = f"{yolo_class_id} {cx:.6f} {cy:.6f} {w_norm:.6f} {h_norm:.6f}"
        composite_image = Image.fromarray(_io.to_host(composite), "RGB")
    except ValueError as ve:
        print(f"Erreur de valeur [{overlay_path.name} + {background_path.name}]: {ve}")
        return None
    except Exception as e:
        print(f"Erreur [{overlay_path.name} + {background_path.name}]: Échec pendant le processus de superposition: {e}")
        return None

    saved: List[Path] = []
    img_out = image_target_dir / f"{overlay_path.stem}{background_path.suffix}"
    label_out = label_target_dir / f"{overlay_path.stem}.txt"
    try:
        composite_image.save(img_out)
        saved.append(img_out)
        with open(label_out, "w", encoding="utf-8") as f:
            f.write(yolo_label_str)
        saved.append(label_out)
        return saved
    except Exception as e_save:
        print(f"Erreur [{overlay_path.name} + {background_path.name}]: Échec lors de la sauvegarde: {e_save}")
        for p in saved:
            try:
                if p.exists():
                    p.unlink()
            except OSError:
                print(f"Avertissement: Impossible de nettoyer le fichier partiellement créé {p}")
        return None


# the reference keeps a deprecated twin with the same behaviour (overlays.py:190-354)
process_overlay_pair = paste_overlay_onto_background
paste_overlay_onto_background.uses_gpu = True
