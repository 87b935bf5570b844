"""change_label_class — drop-in for reference transforms/labels.py:7-65 (host-only text work)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Optional


def change_label_class(input_path: Path, output_dirs: List[Path], cls_mapping: Dict[int, int] = {0: 0},
                       **options: Any) -> Optional[Path]:
    """Rewrites the class id (first token) of every non-blank YOLO label line through
    ``cls_mapping`` (ids absent from the mapping stay), tokens re-joined by single spaces.
    Any failure prints, removes the partial output and returns None."""
    output_path = output_dirs[0] / input_path.name
    try:
        with input_path.open("r", encoding="utf-8") as src, output_path.open("w", encoding="utf-8") as dst:
            for line in src:
                tokens = line.strip().split()
                if not tokens:
                    continue
                cls_id = int(tokens[0])
                tokens[0] = str(cls_mapping.get(cls_id, cls_id))
                dst.write(" ".join(tokens) + "\n")
        return output_path
    except Exception as e:
        print(f"Problème : {e}")
        if output_path.exists():
            output_path.unlink()
        return None
