"""ctypes binding of libippgpu (include/ippgpu.h) — the C-ABI drop-in boundary.

There is deliberately no fallback: if the shared library is missing or a call
fails, an exception is raised.  Error codes map to the exception class the
reference's transform would raise for the same mistake.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_NAME = "libippgpu.so"
_lib = None


class IppError(RuntimeError):
    """A CUDA runtime error reported by libippgpu (positive return code)."""


def lib_path() -> Path:
    env = os.environ.get("IPPGPU_LIB")
    return Path(env) if env else Path(__file__).resolve().parent / _LIB_NAME


class HsvRange(C.Structure):
    _fields_ = [("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3), ("zone", C.c_int32 * 4)]


IPP_MAX_HSV_RANGES = 16


class ChainGeom(C.Structure):
    _fields_ = [("frame_w", C.c_int32), ("frame_h", C.c_int32),
                ("crop_l", C.c_int32), ("crop_t", C.c_int32), ("crop_w", C.c_int32), ("crop_h", C.c_int32),
                ("bg_w", C.c_int32), ("bg_h", C.c_int32),
                ("n_ranges", C.c_int32),
                ("ranges", HsvRange * IPP_MAX_HSV_RANGES)]


class PlanParams(C.Structure):
    _fields_ = [("num_rotations", C.c_int32), ("include_original", C.c_int32),
                ("angle_min", C.c_double), ("angle_max", C.c_double),
                ("pool", C.c_int32 * 4), ("pool_len", C.c_int32),
                ("choose_random", C.c_int32), ("sym_include_original", C.c_int32),
                ("yolo_class_id", C.c_int32),
                ("scale_min", C.c_double), ("scale_max", C.c_double),
                ("bg_w", C.c_int32), ("bg_h", C.c_int32), ("n_bgs", C.c_int32), ("_pad", C.c_int32)]


class PlanTotals(C.Structure):
    _fields_ = [("table_ints", C.c_int64), ("table_rows", C.c_int64), ("sprite_ws_bytes", C.c_int64),
                ("band_units", C.c_int64),
                ("gather_tiles", C.c_int32), ("h_tiles", C.c_int32), ("v_tiles", C.c_int32), ("labels_bytes", C.c_int32),
                ("max_rows", C.c_int32), ("max_band_warps", C.c_int32), ("fused_smem_bytes", C.c_int32), ("vstaged_smem_bytes", C.c_int32)]


_u8p = C.c_void_p      # device pointers travel as integers
_i32p = C.c_void_p
_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

# name -> (restype, argtypes); mirrors include/ippgpu.h declaration by declaration
PROTOTYPES = {
    "ipp_version": (_int, []),
    "ipp_last_error_string": (C.c_char_p, []),
    "ipp_sm_count": (_int, [C.POINTER(C.c_int)]),
    "ipp_launch_count": (C.c_longlong, []),
    "ipp_timing_enable": (_int, [_int]),
    "ipp_timing_report": (_int, [C.c_char_p, _int]),
    "ipp_timing_timeline": (_int, [C.c_char_p, _int]),
    "ipp_remap_u8": (_int, [_u8p, _i64, _int, _int, _int, _int, _int, _int, _int, _int, _int,
                            _u8p, _i64, _int, _i64, _i64, _vp]),
    "ipp_symmetries_u8": (_int, [_u8p, _i64, _int, _int, _int, _u8p, _u8p, _u8p, _u8p, _i64, _int, _i64, _i64, _vp]),
    "ipp_to_rgba_u8": (_int, [_u8p, _i64, _int, _int, _int, _int, _u8p, _i64, _vp]),
    "ipp_rot90_rgba_u8": (_int, [_u8p, _i64, _int, _int, _int, _u8p, _i64, _vp]),
    "ipp_alpha_bbox_u8": (_int, [_u8p, _i64, _int, _int, _int, _int, _i32p, _vp]),
    "ipp_bbox_reset": (_int, [_i32p, _int, _vp]),
    "ipp_rotate_rgba_u8": (_int, [_u8p, _i64, _int, _int, _int, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                  _u8p, _i64, _int, _int, _i32p, _vp]),
    "ipp_resample_rgba_u8": (_int, [_u8p, _i64, _int, _int, _u8p, _i64, _int, _int,
                                    _i32p, _i32p, _int, _i32p, _i32p, _int, _u8p, _vp]),
    "ipp_lanczos_ksize": (_int, [_int, _int]),
    "ipp_lanczos_coeffs_host": (_int, [_int, _int, _vp, _vp]),
    "ipp_paste_rgba_on_rgb_u8": (_int, [_u8p, _i64, _int, _int, _u8p, _i64, _int, _int, _int, _int,
                                        _int, _int, _int, _int, _u8p, _i64, _vp]),
    "ipp_hsv_ranges_to_alpha_u8": (_int, [_u8p, _i64, _int, _int, C.POINTER(HsvRange), _int, _u8p, _i64,
                                          _int, _i64, _i64, _vp]),
    "ipp_keep_largest_workspace_bytes": (_i64, [_int, _int, _int]),
    "ipp_keep_largest_cc_u8": (_int, [_u8p, _i64, _int, _int, _i32p, _vp, _i64, _int, _i64, _vp]),
    "ipp_boxes_square_crop_f32": (_int, [_vp, _i32p, _i32p, _vp, _i32p, _int, _vp]),
    "ipp_boxes_xyxy2xywhn_f32": (_int, [_i32p, _i32p, _vp, _int, _vp]),
    "ipp_plan_rotations_host": (_int, [_vp, _int, C.POINTER(PlanParams), _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp]),
    "ipp_plan_items_host": (_int, [_vp, _vp, _int, _vp, C.POINTER(PlanParams), _vp, _vp, _vp, _vp, _int, _vp,
                                   _vp, _int, C.POINTER(PlanTotals)]),
    "ipp_lanczos_row_host": (_int, [_int, _int, _int, _vp, _vp]),
    "ipp_chain_tables": (_int, [_vp, _int, _int, _i32p, _i32p, _int, C.c_double, _vp]),
    "ipp_chain_mask_workspace_bytes": (_i64, [C.POINTER(ChainGeom), _int]),
    "ipp_chain_stage_a": (_int, [_u8p, _i64, _i64, C.POINTER(ChainGeom), _i32p, _vp, _i64, _int, _vp]),
    "ipp_chain_stage_b1": (_int, [C.POINTER(ChainGeom), _vp, _int, _int, _vp, _i32p, _vp]),
    "ipp_chain_stage_b2": (_int, [_u8p, _i64, _i64, C.POINTER(ChainGeom), _vp, _int, _int, _int, _int,
                                  _i32p, _vp, _int, _int, _int, _vp, _u8p, _vp]),
    "ipp_chain_layout_host": (_int, [_vp, _int, C.POINTER(PlanTotals)]),
    "ipp_host_round15": (C.c_double, [C.c_double]),
    "ipp_host_hypot": (C.c_double, [C.c_double, C.c_double]),
    "ipp_host_overlay_size": (_int, [_int, _int, _int, _int, C.c_double, _vp]),
    "ipp_host_format6": (_int, [C.c_double, C.c_char_p, C.POINTER(C.c_double)]),
    "ipp_chain_stage_c": (_int, [_u8p, _i64, _i64, C.POINTER(ChainGeom), _vp, _int, _u8p, _u8p, _i64, _i64, _int, _vp]),
    "ipp_chain_upload_windows": (_int, [_vp, _i64, _i64, _int, _int, _int, _vp, _int, _int, _u8p, _vp]),
    "ipp_crop_batch_u8": (_int, [_u8p, _i64, _i64, _int, _int, _int, _i32p, _int, _int, _u8p, _i64, _i64, _int, _vp]),
    "ipp_plan_rotate_batch_host": (_int, [_int, _int, _int, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "ipp_rotate_batch_u8": (_int, [_u8p, _i64, _i64, _int, _int, _int, _int, _vp, _int, _int, _int, _u8p, _i32p, _vp]),
    "ipp_plan_overlay_batch_host": (_int, [_int, _int, _int, _int, C.POINTER(PlanParams), _vp, _vp, _vp, _vp, _vp, _int,
                                           C.POINTER(PlanTotals)]),
    "ipp_chain_layout_planar_host": (_int, [_vp, _int, C.POINTER(PlanTotals)]),
    "ipp_sprite_planes_u8": (_int, [_u8p, _i64, _i64, _int, _int, _int, _vp, _int, _u8p, _vp]),
    "ipp_chain_stage_c_rect": (_int, [_u8p, _i64, _i64, _int, _int, _vp, _int, _u8p, _u8p, _i64, _i64, _int, _int, _vp]),
    "ipp_sprites_export_u8": (_int, [_vp, _int, _u8p, _u8p, _vp, _vp]),
    "ipp_plan_enhance_host": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ipp_enhance_workspace_bytes": (_i64, [_int, _int, _int, _int]),
    "ipp_enhance_batch_u8": (_int, [_u8p, _u8p, _int, _int, _int, _vp, _u8p, _i32p, _int, _vp, _i64, _vp]),
    "ipp_bgra_from_keep_u8": (_int, [_u8p, _i64, _i64, C.POINTER(ChainGeom), _vp, _u8p, _i64, _i64, _int, _vp]),
}

EXCEPTION_FOR_CODE = {-1: ValueError, -2: ValueError, -3: MemoryError, -4: NotImplementedError}


def load():
    """Load libippgpu.so and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.is_file():
        raise ImportError(
            f"{path} not found: build it with `python -m image_processor_pipeline_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = ""):
    if code == 0:
        return
    msg = load().ipp_last_error_string().decode("utf-8", "replace")
    if code < 0:
        raise EXCEPTION_FOR_CODE.get(code, RuntimeError)(msg or what)
    raise IppError(f"{what}: CUDA error {code}: {msg}")


def launch_count() -> int:
    return int(load().ipp_launch_count())


def timing_enable(on: bool):
    check(load().ipp_timing_enable(int(on)), "ipp_timing_enable")


def timing_report() -> dict:
    """{kernel: (launches, total_ms)} since the previous report (synchronises)."""
    buf = C.create_string_buffer(1 << 16)
    check(load().ipp_timing_report(buf, len(buf)), "ipp_timing_report")
    out = {}
    for part in buf.value.decode().split(";"):
        if part:
            name, cnt, ms = part.rsplit(":", 2)
            out[name] = (int(cnt), float(ms))
    return out


def timing_timeline() -> list:
    """[(kernel, stream_id, start_ms, end_ms)] per launch since the previous report, in launch order (synchronises)."""
    buf = C.create_string_buffer(1 << 22)
    check(load().ipp_timing_timeline(buf, len(buf)), "ipp_timing_timeline")
    out = []
    for part in buf.value.decode().split(";"):
        if part:
            name, sid, t0, t1 = part.rsplit(":", 3)
            out.append((name, int(sid), float(t0), float(t1)))
    return out
