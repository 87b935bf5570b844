"""ultralytics.utils.ops box converters (reference crop_square.py:8, overlays.py:8),
restated from the published upstream source: float32 result, input-dtype math."""
import numpy as np


def xywhn2xyxy(x, w=640, h=640, padw=0, padh=0):
    y = np.empty_like(x, dtype=np.float32)
    xc, yc, xw, xh = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    half_w, half_h = xw / 2, xh / 2
    y[..., 0] = w * (xc - half_w) + padw
    y[..., 1] = h * (yc - half_h) + padh
    y[..., 2] = w * (xc + half_w) + padw
    y[..., 3] = h * (yc + half_h) + padh
    return y


def xyxy2xywhn(x, w=640, h=640, clip=False, eps=0.0):
    y = np.empty_like(x, dtype=np.float32)
    x1, y1, x2, y2 = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    y[..., 0] = ((x1 + x2) / 2) / w
    y[..., 1] = ((y1 + y2) / 2) / h
    y[..., 2] = (x2 - x1) / w
    y[..., 3] = (y2 - y1) / h
    return y
