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


def _pil_coverage(text, size, path=None, antialias=True):
    from PIL import Image, ImageDraw, ImageFont
    if path is None or not os.path.exists(path):
        path = next((p for p in _FONT_SEARCH if p and os.path.exists(p)), None)
    if path is None:
        raise RuntimeError("no TrueType file available for the PIL coverage provider")
    f = ImageFont.truetype(path, size)
    ascent, descent = f.getmetrics()
    width = int(np.ceil(f.getlength(text)))
    img = Image.new("L", (max(width, 1), ascent + descent), 0)
    d = ImageDraw.Draw(img)
    if not antialias:
        d.fontmode = "1"          # FreeType mono rendering, like SDL_ttf's TTF_RenderText_Solid
    d.text((0, 0), text, font=f, fill=255)  # anchor 'la': top = ascender
    return np.asarray(img, dtype=np.uint8)


def _pil_advance(ch, size, path=None):
    from PIL import ImageFont
    if path is None or not os.path.exists(path):
        path = next((p for p in _FONT_SEARCH if p and os.path.exists(p)), None)
    return ImageFont.truetype(path, size).getlength(ch)


class Font(object):
    def __init__(self, name, size):
        self._name, self._size = name, int(size)

    def render(self, text, antialias, color, background=None):
        from . import Surface
        if _provider is not None:
            cov = _provider(text, self._size)
            if not antialias:
                cov = np.where(cov >= 128, 255, 0).astype(np.uint8)
        elif antialias:
            cov = _pil_coverage(text, self._size, self._name if isinstance(self._name, str) else None, True)
        else:
            # SDL_ttf's solid renderer places glyph bitmaps one by one at integer pen positions (advance
            # rounded to whole pixels); tools/build_car_glyphs.py builds the device glyph atlas the same way
            path = self._name if isinstance(self._name, str) else None
            parts, pen = [], 0
            for ch in text:
                g = _pil_coverage(ch, self._size, path, False)
                adv = int(round(_pil_advance(ch, self._size, path)))
                parts.append((pen, g))
                pen += adv
            height = parts[0][1].shape[0] if parts else 1
            width = max([p + g.shape[1] for p, g in parts] + [1])
            cov = np.zeros((height, width), np.uint8)
            for p0, g in parts:
                cov[:, p0:p0 + g.shape[1]] = np.maximum(cov[:, p0:p0 + g.shape[1]], np.where(g > 0, 255, 0).astype(np.uint8))
        h, w = cov.shape
        s = Surface((w, h))
        s.rgb[:, :] = color[:3]
        s.alpha = cov.copy()
        s.colorkey = not antialias
        return s
