"""enhance_image — drop-in for reference transforms/tranfo.py:9-55 (photometric jitter, the consumer of the
'sample' pairing mode, pipeline.py:209-225)."""
from __future__ import annotations

import random
from pathlib import Path
from typing import Any, List, Optional

import numpy as np
from PIL import Image

from .. import _io, batch
from ..utils.utils import _validate_dirs


def enhance_image(input_image: Path, apply_blur: bool, apply_rgb: bool, output_dirs: List[Path], **options: Any) -> Optional[Path]:
    """Brightness / Contrast / Color jitter (three ``uniform(0.7, 1.3)``), GaussianBlur(``uniform(0.5, 3)``) when
    ``apply_blur``, per-channel ``point`` tables (256 ``uniform(0.75, 1.25)`` per channel) when ``apply_rgb`` — same draws in
    the same order from the module-level ``random`` (tranfo.py:38-51), same output name (``<output_dirs[0]>/<input name>``)
    and format (by extension).  The arithmetic runs on the GPU (ipp_enhance_batch_u8), bit-exact against Pillow."""
    destination_img = _validate_dirs(output_dirs, 1)
    output_path = destination_img / input_image.name
    with Image.open(input_image) as im:
        rgb = np.asarray(im.convert("RGB"))
    out = batch.enhance_batch(_io.to_device(rgb).unsqueeze(0), [bool(apply_blur)], [bool(apply_rgb)], random)
    Image.fromarray(_io.to_host(out[0]), "RGB").save(output_path)
    return output_path


enhance_image.uses_gpu = True
