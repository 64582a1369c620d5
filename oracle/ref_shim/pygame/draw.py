"""pygame.draw stand-in (oracle side): clipped solid rectangle fill only."""


def rect(surface, color, rect, width=0):
    assert width == 0, "only filled rectangles are used by the reference Pong path"
    x, y, w, h = rect
    x0, y0 = max(0, x), max(0, y)
    x1, y1 = min(surface._w, x + w), min(surface._h, y + h)
    if x1 > x0 and y1 > y0:
        surface.rgb[y0:y1, x0:x1] = color[:3]
    return rect
