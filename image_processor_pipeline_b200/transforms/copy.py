"""copy_img_with_labels / copy_files — drop-in for reference transforms/copy.py:6-67 (file plumbing between steps: no pixels,
no GPU work; kept so that a workflow written against the reference imports unchanged)."""
from __future__ import annotations

import shutil
from pathlib import Path
from typing import Any, List, Optional, Tuple

from ..utils.utils import _validate_dirs


def copy_img_with_labels(input_image_path: Path, input_label_path: Path, output_dirs: List[Path], **options: Any) -> Optional[List[Path]]:
    """Copies an image and its label file into output_dirs[0] / output_dirs[1] (copy.py:6-30): ValueError with fewer than
    two output directories, ``None`` (after a message) when a copy fails."""
    if len(output_dirs) < 2:
        raise ValueError(f"Pas assez de dossiers de sortie. {output_dirs}")
    try:
        img_out = shutil.copy2(input_image_path, output_dirs[0])
        lbl_out = shutil.copy2(input_label_path, output_dirs[1])
        return [Path(img_out), Path(lbl_out)]
    except IOError as io:
        print(f"Impossible de copier le fichier : {io}")
        return None
    except Exception as e:
        print(f"Autre erreur : {e}")
        return None


def copy_files(input_file: Path, output_dirs: List[Path], suffix: str = None, replace_params: Tuple[str, str] = None,
               **options: Any) -> Optional[Path]:
    """Copies a file into output_dirs[0], optionally renamed ``<stem>_<suffix><ext>`` or with ``name.replace(*replace_params)``
    (copy.py:32-67); the two renaming options exclude each other (ValueError)."""
    output_dir = _validate_dirs(output_dirs, 1)
    if suffix and replace_params:
        raise ValueError("un seul des 2 paramètres `replace_param` et `suffix` doit être renseigné")
    if suffix:
        out = shutil.copy2(input_file, output_dir / input_file.with_stem(f"{input_file.stem}_{suffix}").name)
    elif replace_params:
        out = shutil.copy2(input_file, output_dir / input_file.name.replace(replace_params[0], replace_params[1]))
    else:
        out = shutil.copy2(input_file, output_dir)
    return Path(out)
