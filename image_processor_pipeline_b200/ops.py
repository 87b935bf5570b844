"""In-memory operator API: the reference's per-image arithmetic on CUDA tensors.

Every function takes/returns ``torch.uint8`` CUDA tensors in HWC layout (the
layout ``cv2.imread`` / ``np.asarray(PIL.Image)`` produce) and calls the
hand-written sm_100a kernels through the C ABI of ``include/ippgpu.h``.
PyTorch only owns the device buffers and the stream.  No CPU fallback exists:
a CPU tensor or a missing library raises.

Host-side scalar math that the reference does in Python (rotation matrix,
overlay size, inRange bound rounding, box bookkeeping) is restated in
``hostmath.py``; random draws stay with the caller.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native, hostmath

INT_MAX = 2 ** 31 - 1


def _lib():
    return _native.load()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, name: str = "image"):
    if not isinstance(t, torch.Tensor) or t.dtype != torch.uint8:
        raise TypeError(f"{name} must be a torch.uint8 tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (there is no CPU fallback)")
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    if t.dim() != 3:
        raise ValueError(f"{name} must be HW or HWC")
    # pixels must be dense inside a row for EVERY channel count (a column-strided gray view gray[:, ::2] is
    # unsqueezed to HWC with stride(1) != 1): the kernels take a row pitch only
    if t.stride(1) != t.shape[2] or (t.shape[2] > 1 and t.stride(2) != 1):
        t = t.contiguous()
    return t


def _geom(t: torch.Tensor):
    """(ptr, pitch, w, h, c) of an HWC tensor whose rows are dense."""
    h, w, c = t.shape
    return t.data_ptr(), t.stride(0) if h > 1 else w * c, w, h, c


# --------------------------------------------------------------------------- flip / crop
def flip_crop(img: torch.Tensor, x0: int = 0, y0: int = 0, w: Optional[int] = None, h: Optional[int] = None,
              flip_h: bool = False, flip_v: bool = False) -> torch.Tensor:
    """out[y, x] = img[y0 + (flip_v ? h-1-y : y), x0 + (flip_h ? w-1-x : x)] — cv2.flip /
    slicing / PIL crop (reference symmetry.py:114-119, recadrages.py:46, crop_square.py:196)."""
    squeeze = img.dim() == 2
    t = _req(img)
    ptr, pitch, W, H, c = _geom(t)
    w = W - x0 if w is None else w
    h = H - y0 if h is None else h
    out = torch.empty((h, w, c), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        _native.check(_lib().ipp_remap_u8(ptr, pitch, W, H, c, x0, y0, w, h, int(flip_h), int(flip_v),
                                          out.data_ptr(), w * c, 1, 0, 0, _stream()), "ipp_remap_u8")
    return out.squeeze(-1) if squeeze else out


def symmetries(img: torch.Tensor, syms: Sequence[str]) -> List[torch.Tensor]:
    """One read of ``img``, one write per requested variant in {'o','h','v','hv'}
    (reference symmetry.py:114-131)."""
    squeeze = img.dim() == 2
    t = _req(img)
    ptr, pitch, W, H, c = _geom(t)
    slot = {"o": 0, "h": 1, "v": 2, "hv": 3}
    for s in syms:
        if s not in slot:
            raise ValueError(f"`pool` contient des éléments invalides : {[s]}")
    bufs = {}
    for s in syms:
        if s not in bufs:
            bufs[s] = torch.empty((H, W, c), dtype=torch.uint8, device=t.device)
    ptrs = [bufs[k].data_ptr() if k in bufs else None for k in ("o", "h", "v", "hv")]
    with torch.cuda.device(t.device):
        _native.check(_lib().ipp_symmetries_u8(ptr, pitch, W, H, c, *ptrs, W * c, 1, 0, 0, _stream()),
                      "ipp_symmetries_u8")
    out = [bufs[s] for s in syms]
    return [o.squeeze(-1) for o in out] if squeeze else out


def to_rgba(img: torch.Tensor, swap_rb: bool = False) -> torch.Tensor:
    """L / RGB / RGBA -> RGBA with A=255 (PIL convert('RGBA'), reference rotations.py:55);
    ``swap_rb`` converts between cv2's BGR(A) and PIL's RGB(A) byte order."""
    t = _req(img)
    ptr, pitch, W, H, c = _geom(t)
    out = torch.empty((H, W, 4), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        _native.check(_lib().ipp_to_rgba_u8(ptr, pitch, W, H, c, int(swap_rb), out.data_ptr(), W * 4, _stream()),
                      "ipp_to_rgba_u8")
    return out


# --------------------------------------------------------------------------- bbox
def _bbox_from_acc(acc) -> Optional[Tuple[int, int, int, int]]:
    x0, y0, x1, y1 = (int(v) for v in acc)
    if x1 <= 0:
        return None
    return x0, y0, x1, y1


def nonzero_bbox(img: torch.Tensor, alpha_only: Optional[bool] = None) -> Optional[Tuple[int, int, int, int]]:
    """PIL getbbox(): (x0, y0, x1, y1) exclusive or None (reference rotations.py:99,
    recadrages.py:73).  alpha_only defaults to True for 2/4-channel images."""
    t = _req(img)
    ptr, pitch, W, H, c = _geom(t)
    if alpha_only is None:
        alpha_only = c in (2, 4)
    acc = torch.empty(4, dtype=torch.int32, device=t.device)
    with torch.cuda.device(t.device):
        _native.check(_lib().ipp_alpha_bbox_u8(ptr, pitch, W, H, c, int(alpha_only), acc.data_ptr(), _stream()),
                      "ipp_alpha_bbox_u8")
    return _bbox_from_acc(acc.cpu().tolist())


# --------------------------------------------------------------------------- rotation
def rotate_expand(rgba: torch.Tensor, angle: float, resample: str = "nearest", with_bbox: bool = False):
    """``Image.rotate(angle, resample, expand=True)`` on an RGBA tensor (reference
    rotations.py:96).  Returns the canvas, or (canvas, bbox) when with_bbox (bbox =
    PIL getbbox() of the canvas, reduced in the same launch)."""
    t = _req(rgba, "rgba")
    if t.shape[2] != 4:
        raise ValueError("rotate_expand expects an RGBA image")
    ptr, pitch, W, H, _ = _geom(t)
    plan = hostmath.rotation_plan(W, H, angle)
    acc = None
    if plan.kind != "affine":
        # PIL fast paths (Image.py:2509-2521): copy / transpose, no resampling
        if plan.kind == "copy":
            out = t.clone()
        elif plan.kind == "rot180":
            out = flip_crop(t, flip_h=True, flip_v=True)
        else:
            out = torch.empty((W, H, 4), dtype=torch.uint8, device=t.device)
            with torch.cuda.device(t.device):
                _native.check(_lib().ipp_rot90_rgba_u8(ptr, pitch, W, H, int(plan.kind == "rot90"),
                                                       out.data_ptr(), H * 4, _stream()), "ipp_rot90_rgba_u8")
        if not with_bbox:
            return out
        return out, nonzero_bbox(out, alpha_only=True)
    nw, nh = plan.size
    out = torch.empty((nh, nw, 4), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        if with_bbox:
            acc = torch.empty(4, dtype=torch.int32, device=t.device)
            _native.check(_lib().ipp_bbox_reset(acc.data_ptr(), 1, _stream()), "ipp_bbox_reset")
        mode = {"nearest": 0, "bilinear": 1}[resample]
        fixed = (C.c_int32 * 6)(*plan.fixed) if mode == 0 else None
        f64 = (C.c_double * 6)(*plan.matrix) if mode == 1 else None
        if mode == 0 and not plan.fixed_ok:
            raise NotImplementedError("rotation outside the 16.16 fixed-point range of libImaging affine_fixed")
        _native.check(_lib().ipp_rotate_rgba_u8(ptr, pitch, W, H, mode, fixed, f64, out.data_ptr(), nw * 4, nw, nh,
                                                acc.data_ptr() if acc is not None else None, _stream()),
                      "ipp_rotate_rgba_u8")
    if not with_bbox:
        return out
    return out, _bbox_from_acc(acc.cpu().tolist())


def rotate_and_fit(rgba: torch.Tensor, angle: float, resample: str = "nearest"):
    """rotate + getbbox + crop (reference rotations.py:96-110).  Returns (image, bbox);
    bbox None => the un-cropped canvas is returned, as the reference does."""
    canvas, bbox = rotate_expand(rgba, angle, resample, with_bbox=True)
    if bbox is None:
        return canvas, None
    x0, y0, x1, y1 = bbox
    if (x0, y0, x1, y1) == (0, 0, canvas.shape[1], canvas.shape[0]):
        return canvas, bbox
    return flip_crop(canvas, x0, y0, x1 - x0, y1 - y0), bbox


# --------------------------------------------------------------------------- resample / paste
def _tables(in_size: int, out_size: int, device):
    lib = _lib()
    ksize = lib.ipp_lanczos_ksize(in_size, out_size)
    bounds = np.empty((out_size, 2), dtype=np.int32)
    kk = np.empty((out_size, ksize), dtype=np.int32)
    _native.check(lib.ipp_lanczos_coeffs_host(in_size, out_size, bounds.ctypes.data, kk.ctypes.data),
                  "ipp_lanczos_coeffs_host")
    return ksize, torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device)


def resize_lanczos(rgba: torch.Tensor, size: Tuple[int, int], scalar_kernels: bool = False) -> torch.Tensor:
    """``Image.resize(size, LANCZOS)`` on RGBA (reference overlays.py:129); bit-exact.  Runs the tensor-core passes of
    the batched path (batch.resize_lanczos_batch with one sprite); ``scalar_kernels=True`` keeps the two scalar
    integer passes of resample.cu (ipp_resample_rgba_u8, caller-provided tables)."""
    t = _req(rgba, "rgba")
    if t.shape[2] != 4:
        raise ValueError("resize_lanczos expects an RGBA image")
    nw, nh = int(size[0]), int(size[1])
    if nw <= 0 or nh <= 0:
        raise ValueError("height and width must be > 0")          # PIL's message for empty sizes
    if not scalar_kernels:
        from . import batch
        return batch.resize_lanczos_batch(t.contiguous().unsqueeze(0), [(nw, nh)])[0]
    ptr, pitch, W, H, _ = _geom(t)
    dev = t.device
    out = torch.empty((nh, nw, 4), dtype=torch.uint8, device=dev)
    hk = vk = 0
    hb = hkk = vb = vkk = tmp = None
    if nw != W:
        hk, hb, hkk = _tables(W, nw, dev)
    if nh != H:
        vk, vb, vkk = _tables(H, nh, dev)
    if hk and vk:
        tmp = torch.empty((H, nw, 4), dtype=torch.uint8, device=dev)
    p = lambda x: x.data_ptr() if x is not None else None
    with torch.cuda.device(dev):
        _native.check(_lib().ipp_resample_rgba_u8(ptr, pitch, W, H, out.data_ptr(), nw * 4, nw, nh,
                                                  p(hb), p(hkk), hk, p(vb), p(vkk), vk, p(tmp), _stream()),
                      "ipp_resample_rgba_u8")
    return out


def paste(bg_rgb: torch.Tensor, sprite_rgba: torch.Tensor, pos: Tuple[int, int],
          crop: Optional[Tuple[int, int, int, int]] = None) -> torch.Tensor:
    """``bg.copy().paste(sprite, pos, sprite)`` (reference overlays.py:138-139), optionally
    returning only composite[y0:y0+h, x0:x0+w] for crop=(x0, y0, w, h) (crop_square.py:196)."""
    b = _req(bg_rgb, "background")
    s = _req(sprite_rgba, "sprite")
    if b.shape[2] != 3 or s.shape[2] != 4:
        raise ValueError("paste expects an RGB background and an RGBA sprite")
    bptr, bpitch, W, H, _ = _geom(b)
    sptr, spitch, sw, sh, _ = _geom(s)
    x0, y0, w, h = crop if crop is not None else (0, 0, W, H)
    out = torch.empty((h, w, 3), dtype=torch.uint8, device=b.device)
    with torch.cuda.device(b.device):
        _native.check(_lib().ipp_paste_rgba_on_rgb_u8(bptr, bpitch, W, H, sptr, spitch, sw, sh, int(pos[0]), int(pos[1]),
                                                      x0, y0, w, h, out.data_ptr(), w * 3, _stream()),
                      "ipp_paste_rgba_on_rgb_u8")
    return out


# --------------------------------------------------------------------------- HSV keying
def color_masks_to_alpha(bgr: torch.Tensor, color_ranges, zones=None, use_gimp_scale: bool = False) -> torch.Tensor:
    """BGR (HW3 or NHW3) -> BGRA with alpha = ~OR_k(inRange_k & zone_k)
    (reference filtres_liste.py:90-134)."""
    ranges = hostmath.pack_hsv_ranges(color_ranges, zones, use_gimp_scale)
    batched = bgr.dim() == 4
    t = bgr if batched else bgr.unsqueeze(0)
    if t.dtype != torch.uint8 or not t.is_cuda:
        raise RuntimeError("bgr must be a CUDA uint8 tensor (there is no CPU fallback)")
    t = t.contiguous()
    n, H, W, c = t.shape
    if c != 3:
        raise ValueError("color_masks_to_alpha expects a 3-channel BGR image")
    out = torch.empty((n, H, W, 4), dtype=torch.uint8, device=t.device)
    arr = (_native.HsvRange * len(ranges))(*ranges)
    with torch.cuda.device(t.device):
        _native.check(_lib().ipp_hsv_ranges_to_alpha_u8(t.data_ptr(), W * 3, W, H, arr, len(ranges), out.data_ptr(), W * 4,
                                                        n, H * W * 3, H * W * 4, _stream()),
                      "ipp_hsv_ranges_to_alpha_u8")
    return out if batched else out[0]


# --------------------------------------------------------------------------- largest component
def keep_largest_component(bgra: torch.Tensor, inplace: bool = False):
    """alpha[not in largest 8-connected component of alpha>1] = 0 (reference
    pixels_isolés.py:32-55).  Accepts HW4 or NHW4.  Returns (cleaned, info) with info an
    int32 CPU array [n, 8] = {x, y, w, h of alpha != 0, area, n_fg, 0, 0}."""
    batched = bgra.dim() == 4
    t = bgra if batched else bgra.unsqueeze(0)
    if t.dtype != torch.uint8 or not t.is_cuda:
        raise RuntimeError("bgra must be a CUDA uint8 tensor (there is no CPU fallback)")
    if t.shape[-1] != 4:
        raise AttributeError("L'image ne contient pas de canal alpha, elle sera ignorée.")
    t = t.contiguous() if inplace else t.clone(memory_format=torch.contiguous_format)
    n, H, W, _ = t.shape
    lib = _lib()
    ws_bytes = lib.ipp_keep_largest_workspace_bytes(W, H, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=t.device)
    info = torch.empty((n, 8), dtype=torch.int32, device=t.device)
    with torch.cuda.device(t.device):
        _native.check(lib.ipp_keep_largest_cc_u8(t.data_ptr(), W * 4, W, H, info.data_ptr(), ws.data_ptr(), ws_bytes,
                                                 n, H * W * 4, _stream()), "ipp_keep_largest_cc_u8")
    info_h = info.cpu().numpy()
    return (t if batched else t[0]), info_h


# --------------------------------------------------------------------------- boxes
def square_crop_boxes(boxes_xywhn: np.ndarray, img_w: int, img_h: int, x0: int, y0: int, crop_size: int, device):
    """Box re-normalisation of reference crop_square.py:167,201-217 on the device.
    Returns (boxes float32 [N,4], valid bool [N]) as numpy arrays."""
    b = np.ascontiguousarray(boxes_xywhn, dtype=np.float64).reshape(-1, 4)
    n = b.shape[0]
    d_in = torch.from_numpy(b).to(device)
    offs = torch.tensor([0, n], dtype=torch.int32, device=device)
    params = torch.tensor([img_w, img_h, x0, y0, crop_size], dtype=torch.int32, device=device)
    out = torch.empty((n, 4), dtype=torch.float32, device=device)
    valid = torch.empty(n, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _native.check(_lib().ipp_boxes_square_crop_f32(d_in.data_ptr(), offs.data_ptr(), params.data_ptr(),
                                                       out.data_ptr(), valid.data_ptr(), 1, _stream()),
                      "ipp_boxes_square_crop_f32")
    return out.cpu().numpy(), valid.cpu().numpy().astype(bool)


def xyxy2xywhn(xyxy: np.ndarray, img_w: int, img_h: int, device) -> np.ndarray:
    """int xyxy boxes -> float32 xywhn (reference overlays.py:143-146)."""
    b = np.ascontiguousarray(xyxy, dtype=np.int32).reshape(-1, 4)
    n = b.shape[0]
    d_in = torch.from_numpy(b).to(device)
    wh = torch.tensor([[img_w, img_h]] * n, dtype=torch.int32, device=device)
    out = torch.empty((n, 4), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _native.check(_lib().ipp_boxes_xyxy2xywhn_f32(d_in.data_ptr(), wh.data_ptr(), out.data_ptr(), n, _stream()),
                      "ipp_boxes_xyxy2xywhn_f32")
    return out.cpu().numpy()
