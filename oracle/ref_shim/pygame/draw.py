"""pygame.draw stand-in (oracle side).

rect: pygame 1.9.x implements draw.rect by building the 4 corner points (l, t), (r, t), (r, b), (l, b) with
r = x + w - 1, b = y + h - 1 from the int-truncated Rect and calling polygon() [restated from memory].
polygon: draw_fillpoly of pygame 1.9 draw.c: vertices truncated to int; for every scanline y in
[miny, maxy] collect x = (y - y1) * (x2 - x1) / (y2 - y1) + x1 (C integer division) for the edges with
(y1 <= y < y2) or (y == maxy and y1 < y <= y2), sort, fill between pairs inclusive [restated from memory]."""
import numpy as np


def _cdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def polygon(surface, color, points, width=0):
    assert width == 0
    vx = [int(p[0]) for p in points]
    vy = [int(p[1]) for p in points]
    n = len(vx)
    miny, maxy = min(vy), max(vy)
    col = np.array([int(c) for c in color[:3]], np.uint8)
    H, W = surface.rgb.shape[:2]
    for y in range(max(miny, 0), min(maxy, H - 1) + 1):
        xs = []
        for i in range(n):
            i1 = i - 1 if i else n - 1
            y1, y2, x1, x2 = vy[i1], vy[i], vx[i1], vx[i]
            if y1 < y2:
                pass
            elif y1 > y2:
                y1, y2, x1, x2 = y2, y1, x2, x1
            else:
                continue
            if (y1 <= y < y2) or (y == maxy and y1 < y <= y2):
                xs.append(_cdiv((y - y1) * (x2 - x1), (y2 - y1)) + x1)
        xs.sort()
        for k in range(0, len(xs) - 1, 2):
            a, b = max(xs[k], 0), min(xs[k + 1], W - 1)
            if b >= a:
                surface.rgb[y, a:b + 1] = col
    return None


def rect(surface, color, rect, width=0):
    assert width == 0, "only filled rectangles are used by the reference"
    x, y, w, h = (int(v) for v in rect)
    l, t, r, b = x, y, x + w - 1, y + h - 1
    if hasattr(surface, "_simple_rect") or (w > 0 and h > 0 and isinstance(rect, tuple) is False):
        pass
    if w > 0 and h > 0:      # same pixels as the polygon route, without the per-scanline loop
        x0, y0 = max(0, l), max(0, t)
        x1, y1 = min(surface._w - 1, r), min(surface._h - 1, b)
        if x1 >= x0 and y1 >= y0:
            surface.rgb[y0:y1 + 1, x0:x1 + 1] = [int(c) for c in color[:3]]
        return rect
    polygon(surface, color, [(l, t), (r, t), (r, b), (l, b)], 0)
    return rect


def circle(surface, color, center, radius, width=0):
    raise NotImplementedError
