"""pygame.font stand-in (oracle side).

pygame renders text with SDL_ttf + FreeType; neither is available here, so the
glyph coverage comes from a pluggable provider:

  * default provider: PIL/FreeType rendering of a FreeSansBold TrueType file
    (pygame's bundled default font `freesansbold.ttf` is GNU FreeSansBold; the
    reference ships the same face as competitive_rl/pong/FreeSansBold.ttf).
  * `set_coverage_provider(fn)`: fn(text, size) -> (H, W) uint8 alpha coverage.

`Font.render(text, True, color)` returns a per-pixel-alpha Surface of height
ascent+descent (SDL_ttf TTF_FontHeight), ink placed with its top at the
ascender line, exactly the convention SDL_ttf uses for blended rendering.
"""
import os

import numpy as np

_provider = None
_FONT_SEARCH = [
    os.environ.get("CRL_FREESANSBOLD", ""),
    "/root/reference/competitive_rl/pong/FreeSansBold.ttf",
]


def init():
    pass


def set_coverage_provider(fn):
    global _provider
    _provider = fn


def _pil_coverage(text, size):
    from PIL import Image, ImageDraw, ImageFont
    path = next((p for p in _FONT_SEARCH if p and os.path.exists(p)), None)
    if path is None:
        raise RuntimeError("no FreeSansBold.ttf available for the PIL coverage provider")
    f = ImageFont.truetype(path, size)
    ascent, descent = f.getmetrics()
    width = int(np.ceil(f.getlength(text)))
    img = Image.new("L", (max(width, 1), ascent + descent), 0)
    ImageDraw.Draw(img).text((0, 0), text, font=f, fill=255)  # anchor 'la': top = ascender
    return np.asarray(img, dtype=np.uint8)


class Font(object):
    def __init__(self, name, size):
        self._name, self._size = name, int(size)

    def render(self, text, antialias, color, background=None):
        from . import Surface
        cov = (_provider or _pil_coverage)(text, self._size)
        if not antialias:
            cov = np.where(cov >= 128, 255, 0).astype(np.uint8)
        h, w = cov.shape
        s = Surface((w, h))
        s.rgb[:, :] = color[:3]
        s.alpha = cov.copy()
        return s
