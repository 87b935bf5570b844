"""Boundary helpers shared by the transforms (reference utils/utils.py:7-65)."""
from __future__ import annotations

from pathlib import Path
from typing import List, Tuple, Union


def check_path(folder_name, root=None) -> Path:
    """Absolute paths pass through; relative ones are joined to ``root`` (or '.')
    — reference utils/utils.py:7-37."""
    path = Path(folder_name)
    if path.is_absolute():
        return path
    return (Path(root) if root else Path(".")) / path


def _validate_dirs(output_dirs: List[Path], nb_dirs: int) -> Union[Path, Tuple[Path, ...]]:
    """IndexError when fewer than ``nb_dirs`` directories are given; returns a single Path
    for nb_dirs == 1, else the tuple of ALL directories (reference utils/utils.py:39-65)."""
    if len(output_dirs) < nb_dirs:
        raise IndexError(f"Au moins {nb_dirs} dossiers de sortie requis (images, labels). {len(output_dirs)} fournis.")
    paths = tuple(Path(d) for d in output_dirs)
    return paths[0] if nb_dirs == 1 else paths
