"""The C-ABI library loads and exports every symbol include/ippgpu.h declares; argument
validation and the host-only helpers work without a GPU (no device compute here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from image_processor_pipeline_b200 import _native
from oracle import intmodels as M

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "ippgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in ippgpu.h but not exported"
    missing = [s for s in syms if s not in _native.PROTOTYPES]
    assert not missing, f"ctypes prototypes missing for {missing}"
    assert lib.ipp_version() == 100


def test_error_codes_map_to_reference_exception_classes():
    lib = _native.load()
    rc = lib.ipp_remap_u8(None, 0, 4, 4, 3, 0, 0, 4, 4, 0, 0, None, 0, 1, 0, 0, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _native.check(rc, "ipp_remap_u8")
    assert b"NULL" in lib.ipp_last_error_string()
    rc = lib.ipp_remap_u8(1, 12, 4, 4, 3, 2, 2, 4, 4, 0, 0, 1, 12, 1, 0, 0, None)      # crop outside the image
    assert rc == -2
    rc = lib.ipp_hsv_ranges_to_alpha_u8(1, 12, 4, 4, None, 0, 1, 16, 1, 0, 0, None)
    assert rc == -1
    assert lib.ipp_keep_largest_workspace_bytes(1920, 1080, 1) > 1920 * 1080 // 8
    assert lib.ipp_launch_count() == 0


@pytest.mark.parametrize("sizes", [(80, 48), (50, 75), (200, 31), (64, 64), (33, 20), (512, 300), (7, 1), (3, 20), (604, 297)])
def test_host_lanczos_tables_equal_pillow_model(sizes):
    lib = _native.load()
    n_in, n_out = sizes
    ks = lib.ipp_lanczos_ksize(n_in, n_out)
    bounds = np.empty((n_out, 2), np.int32)
    kk = np.empty((n_out, ks), np.int32)
    assert lib.ipp_lanczos_coeffs_host(n_in, n_out, bounds.ctypes.data, kk.ctypes.data) == 0
    mks, mb, mk = M.lanczos_coeffs(n_in, n_out)
    assert ks == mks and np.array_equal(bounds, mb) and np.array_equal(kk, mk)


def test_product_path_has_no_cpu_fallback():
    import torch
    from image_processor_pipeline_b200 import chain, ops
    cpu_img = torch.zeros((8, 8, 3), dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.flip_crop(cpu_img)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chain.run_chain(cpu_img[None], cpu_img[None], chain.ChainParams(color_ranges_to_exclude_hsv=[(0, 0, 0, 180, 255, 150)]))
    src = (ROOT / "image_processor_pipeline_b200").rglob("*.py")
    for f in src:
        assert "oracle" not in f.read_text().replace("oracle/", ""), f"{f} must not import the oracle"


def test_jpeg_library_exports_every_declared_symbol():
    """include/ippjpeg.h (the nvJPEG codec layer, its own shared library) — symbols only, no device work."""
    import ctypes
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "ippjpeg.h").read_text(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(ippjpeg_[a-z0-9_]+)\s*\(", text)))
    assert len(syms) >= 8
    lib = ctypes.CDLL(str(ROOT / "image_processor_pipeline_b200" / "libippjpeg.so"))
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in ippjpeg.h but not exported"


def test_video_library_exports_every_declared_symbol_and_reports_unavailability():
    """include/ippvideo.h (the NVDEC frame source, its own shared library; libnvcuvid is opened at run time) — symbols, and the
    loud failure where there is no decoder: no device work."""
    import ctypes
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "ippvideo.h").read_text(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(ippvideo_[a-z0-9_]+)\s*\(", text)))
    assert len(syms) >= 6
    lib = ctypes.CDLL(str(ROOT / "image_processor_pipeline_b200" / "libippvideo.so"))
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in ippvideo.h but not exported"
    lib.ippvideo_last_error_string.restype = ctypes.c_char_p
    import torch
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert lib.ippvideo_create(ctypes.byref(h), 10) != 0 and not h.value
        assert lib.ippvideo_last_error_string()
