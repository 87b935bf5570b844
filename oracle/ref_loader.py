"""T1 oracle: import the REAL reference (``/root/reference``) behind four shims.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Only usable in the build
container (the GPU box has no ``/root/reference``); ``tests/golden/make_golden.py``
uses it to generate the committed fixtures, and ``tests/test_reference_live.py``
uses it (skipped when the reference is absent) to diff the drop-in's files
against the reference's.

Shims:
  1. a temp dir holding a symlink ``image_processor_pipeline -> /root/reference``
     (the reference's modules import themselves under that package name);
  2. ``icecream.ic`` stub;  3. ``ultralytics`` stub (format sets + box converters);
  4. ``tqdm.notebook.tqdm`` aliased to ``tqdm.tqdm`` (no ipywidgets here).
``crop_square.py`` cannot run as shipped (three one-token defects, SURVEY.md
§8a-4); ``load()`` imports a repaired temp copy of it and reports the repairs.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("IPP_REFERENCE_ROOT", "/root/reference"))
_SHIMS = Path(__file__).resolve().parent / "ref_shims"
_state = {}

CROP_SQUARE_REPAIRS = (
    ("utils._validate_dirs(output_dirs)", "utils._validate_dirs(output_dirs, nb_dirs=2)"),
    ("filepath.isfile()", "filepath.is_file()"),
    ("(clipped[:, 0] < clipped[:, 2]) and (clipped[:, 1] < clipped[:, 3])",
     "(clipped[:, 0] < clipped[:, 2]) & (clipped[:, 1] < clipped[:, 3])"),
)


def available() -> bool:
    return (REFERENCE_ROOT / "pipeline.py").is_file()


def load():
    """Returns a namespace with the reference's modules: .pipeline, .symmetry,
    .rotations, .crop_square (repaired copy), .recadrages, .overlays,
    .pixels_isoles, .filtres_liste, .labels, .video, .utils."""
    if "ns" in _state:
        return _state["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    tmp = Path(tempfile.mkdtemp(prefix="ipp_ref_"))
    (tmp / "image_processor_pipeline").symlink_to(REFERENCE_ROOT, target_is_directory=True)
    sys.path.insert(0, str(tmp))
    sys.path.insert(0, str(_SHIMS))
    import tqdm
    nb = types.ModuleType("tqdm.notebook")
    nb.tqdm = tqdm.tqdm
    sys.modules["tqdm.notebook"] = nb

    ns = types.SimpleNamespace()
    ns.pipeline = importlib.import_module("image_processor_pipeline.pipeline")
    t = "image_processor_pipeline.transforms."
    ns.symmetry = importlib.import_module(t + "symmetry")
    ns.rotations = importlib.import_module(t + "rotations")
    ns.recadrages = importlib.import_module(t + "recadrages")
    ns.overlays = importlib.import_module(t + "overlays")
    ns.pixels_isoles = importlib.import_module(t + "pixels_isolés")
    ns.filtres_liste = importlib.import_module(t + "filtres_liste")
    ns.labels = importlib.import_module(t + "labels")
    ns.video = importlib.import_module(t + "video")
    ns.tranfo = importlib.import_module(t + "tranfo")
    ns.utils = importlib.import_module("image_processor_pipeline.utils.utils")

    src = (REFERENCE_ROOT / "transforms" / "crop_square.py").read_text(encoding="utf-8")
    for old, new in CROP_SQUARE_REPAIRS:
        assert src.count(old) == 1, f"repair anchor not found exactly once: {old!r}"
        src = src.replace(old, new)
    patched = tmp / "crop_square_repaired.py"
    patched.write_text(src, encoding="utf-8")
    spec = importlib.util.spec_from_file_location("crop_square_repaired", patched)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ns.crop_square = mod
    _state["ns"] = ns
    return ns
