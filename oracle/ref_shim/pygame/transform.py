"""pygame.transform stand-in: rotate() restated from pygame 1.9 transform.c (16.16 fixed-point
nearest-neighbour sampling, destination expanded to the rotated bounding box) [from memory]."""
import math

import numpy as np


def scale(surface, size):
    from . import Surface
    return Surface(size)


def rotate(surface, angle):
    from . import Surface
    src = surface.rgb
    sh, sw = src.shape[:2]
    rad = angle * .01745329251994329
    sangle, cangle = math.sin(rad), math.cos(rad)
    x, y = sw, sh
    cx, cy, sx, sy = cangle * x, cangle * y, sangle * x, sangle * y
    nxmax = int(max(abs(cx + sy), abs(cx - sy), abs(-cx + sy), abs(-cx - sy)))
    nymax = int(max(abs(sx + cy), abs(sx - cy), abs(-sx + cy), abs(-sx - cy)))
    dst = Surface((nxmax, nymax))
    dw, dh = nxmax, nymax
    cyi = dh // 2
    xd, yd = (sw - dw) << 15, (sh - dh) << 15
    isin, icos = int(sangle * 65536), int(cangle * 65536)
    ax = (dw << 15) - int(cangle * ((dw - 1) << 15))
    ay = (dh << 15) - int(sangle * ((dw - 1) << 15))
    xmaxval, ymaxval = (sw << 16) - 1, (sh << 16) - 1
    ys = np.arange(dh, dtype=np.int64)[:, None]
    xs = np.arange(dw, dtype=np.int64)[None, :]
    dx = (ax + isin * (cyi - ys)) + xd + icos * xs
    dy = (ay - icos * (cyi - ys)) + yd + isin * xs
    ok = (dx >= 0) & (dy >= 0) & (dx <= xmaxval) & (dy <= ymaxval)
    out = np.zeros((dh, dw, 3), np.uint8)
    out[ok] = src[(dy[ok] >> 16), (dx[ok] >> 16)]
    dst.rgb = out
    return dst
