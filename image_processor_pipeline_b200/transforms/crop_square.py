"""process_square_crop_around_bbox — drop-in for reference transforms/crop_square.py:12-224
with the three one-token defects of the shipped file repaired (SURVEY.md §8a-4:
``_validate_dirs`` argument, ``is_file``, element-wise ``&``)."""
from __future__ import annotations

import random
from pathlib import Path
from typing import Any, List, Optional, Tuple
from warnings import warn

import cv2
import numpy as np

from .. import _io, hostmath, ops
from ..utils import utils


def _load_image(filepath: Path) -> np.ndarray:
    if not filepath.is_file():
        raise FileNotFoundError(f"Image non trouvée: {filepath}")
    img = cv2.imread(str(filepath))
    if img is None:
        raise IOError(f"Impossible de charger l'image {filepath.name} via OpenCV.")
    return img


def _read_bboxes(filepath: Path) -> Tuple[np.ndarray, np.ndarray]:
    if not filepath.is_file():
        raise FileNotFoundError(f"Fichier label non trouvé : {filepath}")
    data = np.loadtxt(filepath, ndmin=2)
    try:
        classes = data[:, 0].astype(int)
        bboxes = data[:, 1:5].astype(float)
    except Exception as e:
        raise ValueError(f"Format invalide dans {filepath.name}: {e}")
    return classes, bboxes


def _save_crop_files(img: np.ndarray, labels: Tuple[np.ndarray, np.ndarray], img_out: Path, label_out: Path) -> None:
    classes, bboxes = labels
    if not cv2.imwrite(str(img_out), img):
        raise IOError(f"Échec écriture de l'image : {img_out}")
    with open(label_out, "w", encoding="utf-8") as f:
        f.write(hostmath.format_label_rows(classes, bboxes))


def process_square_crop_around_bbox(input_image_path: Path, input_label_path: Path, output_dirs: List[Path],
                                    **options: Any) -> Optional[List[Path]]:
    """Square crop of side min(h, w) at a random offset that still contains every box
    (two ``random.randint`` draws), boxes shifted / clipped / re-normalised with the
    reference's float32 semantics; image and label written under the input names.
    Raises IndexError / FileNotFoundError / IOError / ValueError / RuntimeError like the reference."""
    image_target_dir, label_target_dir = utils._validate_dirs(output_dirs, nb_dirs=2)[:2]
    if input_image_path.stem != input_label_path.stem:
        warn(f"Warning [Crop Carré]: image ({input_image_path.name}) et label ({input_label_path.name}) "
             "n'ont pas le même nom. Fichier ignoré et poursuite du traitement...")
    image = _load_image(input_image_path)
    class_ids, bboxes = _read_bboxes(input_label_path)
    height, width = image.shape[:2]
    crop_size, (lbx, ubx), (lby, uby) = hostmath.square_crop_bounds(bboxes, width, height)
    x0 = random.randint(lbx, ubx)
    y0 = random.randint(lby, uby)

    dev = _io.device()
    cropped = ops.flip_crop(_io.to_device(image), x0, y0, crop_size, crop_size)
    if cropped.numel() == 0:
        raise RuntimeError("Le crop a produit une image vide.")
    new_boxes, valid = ops.square_crop_boxes(bboxes, width, height, x0, y0, crop_size, dev)
    if not valid.any():
        raise RuntimeError("Aucune bbox résiduelle après le crop.")
    img_out = image_target_dir / input_image_path.name
    label_out = label_target_dir / input_label_path.name
    _save_crop_files(_io.to_host(cropped), (class_ids[valid], new_boxes[valid]), img_out, label_out)
    return [img_out, label_out]


process_square_crop_around_bbox.uses_gpu = True
