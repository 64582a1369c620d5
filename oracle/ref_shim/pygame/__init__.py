"""Stand-in for pygame 1.9.6 -- TEST INFRASTRUCTURE ONLY (oracle side).

Only the calls that the reference's Pong path makes are restated
(`competitive_rl/pong/base_pong_env.py`): `Rect` (int storage, truncating
setters), `Surface` (numpy RGB), `draw.rect` (clipped fill), `font.Font.render`
(+ alpha `blit`), `surfarray.array3d`, `sprite.Sprite`, `display.quit`.

What is restated here is third-party behaviour that is NOT pinned by any code
in /root/reference (pygame is an un-vendored wheel):
  * pygame 1.9.x `Rect` attribute setters convert with C `(int)` -> truncation
    toward zero (this is what makes vy=-2.7 move -3 px but vy=+2.7 move +2 px).
  * `Rect.centery = y + (h >> 1)`.
  * SDL 1.2 per-pixel-alpha blit: d = d + (((s - d) * a) >> 8)  (arithmetic shift).
"""
import numpy as np

from . import display, draw, font, image, sprite, surfarray, transform  # noqa: F401


def init():
    return (6, 0)


def quit():  # noqa: A001
    pass


def _trunc(v):
    # C `(int)double` semantics: toward zero.
    return int(v)


class Rect(object):
    __slots__ = ("_x", "_y", "_w", "_h")

    def __init__(self, x, y=None, w=None, h=None):
        if y is None:
            x, y, w, h = x
        self._x, self._y, self._w, self._h = _trunc(x), _trunc(y), _trunc(w), _trunc(h)

    def _sx(self, v):
        self._x = _trunc(v)

    def _sy(self, v):
        self._y = _trunc(v)

    x = property(lambda s: s._x, _sx)
    y = property(lambda s: s._y, _sy)
    left = property(lambda s: s._x, _sx)
    top = property(lambda s: s._y, _sy)
    right = property(lambda s: s._x + s._w, lambda s, v: setattr(s, "_x", _trunc(v) - s._w))
    bottom = property(lambda s: s._y + s._h, lambda s, v: setattr(s, "_y", _trunc(v) - s._h))
    width = property(lambda s: s._w)
    height = property(lambda s: s._h)
    w = property(lambda s: s._w)
    h = property(lambda s: s._h)
    centerx = property(lambda s: s._x + (s._w >> 1))
    centery = property(lambda s: s._y + (s._h >> 1))
    center = property(lambda s: (s._x + (s._w >> 1), s._y + (s._h >> 1)))

    def _set_topleft(self, v):
        self._x, self._y = _trunc(v[0]), _trunc(v[1])

    topleft = property(lambda s: (s._x, s._y), _set_topleft)
    size = property(lambda s: (s._w, s._h))

    def __iter__(self):
        return iter((self._x, self._y, self._w, self._h))

    def __repr__(self):
        return "<rect(%d, %d, %d, %d)>" % (self._x, self._y, self._w, self._h)


class Surface(object):
    """RGB (and optionally per-pixel alpha) pixel store, row-major (h, w, 3)."""

    def __init__(self, size, flags=0, depth=0):
        self._w, self._h = int(size[0]), int(size[1])
        self.rgb = np.zeros((self._h, self._w, 3), np.uint8)
        self.alpha = None  # (h, w) uint8 when the surface carries per-pixel alpha

    def get_rect(self):
        return Rect(0, 0, self._w, self._h)

    def get_size(self):
        return (self._w, self._h)

    def get_width(self):
        return self._w

    def get_height(self):
        return self._h

    def fill(self, color, rect=None):
        if rect is None:
            self.rgb[:, :] = [int(c) for c in color[:3]]
        else:
            draw.rect(self, color, rect)

    def subsurface(self, rect):
        x, y, w, h = (int(v) for v in rect)
        assert 0 <= x and 0 <= y and x + w <= self._w and y + h <= self._h, "subsurface rectangle outside surface area"
        sub = Surface.__new__(Surface)
        sub._w, sub._h = w, h
        sub.rgb = self.rgb[y:y + h, x:x + w]
        sub.alpha = None
        return sub

    def blit(self, source, dest):
        if isinstance(dest, Rect):
            dx, dy = dest.x, dest.y
        else:
            dx, dy = int(dest[0]), int(dest[1])
        # clip source rectangle against the destination surface
        sx0, sy0 = max(0, -dx), max(0, -dy)
        sx1 = min(source._w, self._w - dx)
        sy1 = min(source._h, self._h - dy)
        if sx1 <= sx0 or sy1 <= sy0:
            return
        d = self.rgb[dy + sy0:dy + sy1, dx + sx0:dx + sx1].astype(np.int32)
        s = source.rgb[sy0:sy1, sx0:sx1].astype(np.int32)
        if source.alpha is None:
            out = s
        elif getattr(source, "colorkey", False):   # non-antialiased text: 8-bit surface with a colorkey, exact copy
            m = (source.alpha[sy0:sy1, sx0:sx1] > 0)[:, :, None]
            out = np.where(m, s, d)
        else:
            a = source.alpha[sy0:sy1, sx0:sx1].astype(np.int32)[:, :, None]
            out = d + (((s - d) * a) >> 8)  # SDL 1.2 ALPHA_BLEND, arithmetic shift
        self.rgb[dy + sy0:dy + sy1, dx + sx0:dx + sx1] = out.astype(np.uint8)
