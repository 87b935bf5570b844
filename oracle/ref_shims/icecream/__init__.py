"""Stub of `icecream` (absent from the image).  The reference imports ``ic`` at
crop_square.py:9 and overlays.py:10 but never calls it.  TEST INFRASTRUCTURE."""


def ic(*a, **k):
    return a[0] if len(a) == 1 else a
