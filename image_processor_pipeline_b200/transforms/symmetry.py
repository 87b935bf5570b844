"""generate_symmetries — drop-in for reference transforms/symmetry.py:11-149."""
from __future__ import annotations

import random
from pathlib import Path
from typing import Any, List, Optional
from warnings import warn

import cv2

from .. import _io, ops

ALL_SYMS = ("o", "h", "v", "hv")


def generate_symmetries(input_path: Path, output_dirs: List[Path], pool: Optional[List[str]] = None,
                        choose_random: Optional[int] = None, include_original: bool = True,
                        **options: Any) -> Optional[List[Path]]:
    """Writes ``<stem>_<sym><suffix>`` for the selected symmetries among o / h / v / hv.
    Selection, draw (one ``random.sample``), ordering, errors: reference symmetry.py:88-126.
    All variants come from ONE device read of the image (ipp_symmetries_u8)."""
    if not output_dirs:
        raise ValueError(f"Erreur [{input_path.name} - Symétrie]: Aucun dossier de sortie ('output_dirs') fourni.")
    output_dir = output_dirs[0]
    if input_path.suffix.lower()[1:] not in _io.IMG_FORMATS:
        raise ValueError(f"Le fichier {input_path.name} n'est pas un format accepté par Yolo.")
    pool = pool if pool else list(ALL_SYMS)
    bad = [s for s in pool if s not in ALL_SYMS]
    if bad:
        raise ValueError(f"`pool` contient des éléments invalides : {bad}")
    choose_random = len(pool) if choose_random is None else choose_random
    if choose_random > len(pool):
        warn(f"Choix aléatoire de plus d'éléments ({choose_random}) que possible parmi {pool} ({len(pool)}).")
    elif choose_random < 0:
        raise ValueError(f"[{input_path.name} - Symétrie] `choose_random` ({choose_random}) doit être >= 0.")

    image = cv2.imread(str(input_path), cv2.IMREAD_UNCHANGED)
    if image is None:
        raise FileNotFoundError(f"[{input_path.name} - Symétrie] Impossible de charger l'image.")

    selected = random.sample(pool, choose_random)
    if include_original and "o" not in set(selected):
        selected.append("o")

    variants = ops.symmetries(_io.to_device(image), selected) if selected else []
    saved: List[Path] = []
    for sym, dev_img in zip(selected, variants):
        output_path = output_dir / input_path.with_stem(f"{input_path.stem}_{sym}").name
        try:
            if cv2.imwrite(str(output_path), _io.to_host(dev_img, like=image)):
                saved.append(output_path)
            else:
                warn(f"Échec de sauvegarde de la symétrie '{sym}' pour {output_path.name}.")
        except Exception as e_save:  # keep trying the other variants, like the reference
            warn(f"Erreur [{input_path.name} - Symétrie '{sym}']: Échec de sauvegarde : {e_save}.")
    return saved


generate_symmetries.uses_gpu = True
