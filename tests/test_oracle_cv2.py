"""Pins the oracle's restatement of cv2 arithmetic against the real cv2 in the image
(third-party behaviour the reference calls at utils/atari_wrappers.py:216-218)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from pong_oracle import resize_area, warp  # noqa: E402


@pytest.mark.parametrize("dim", [84, 42])
def test_warp_matches_cv2_random(dim):
    rng = np.random.default_rng(dim)
    for _ in range(12):
        rgb = rng.integers(0, 256, (210, 160, 3), dtype=np.uint8)
        ref = cv2.resize(cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY), (dim, dim), interpolation=cv2.INTER_AREA)
        assert np.array_equal(warp(rgb, dim), ref)


@pytest.mark.parametrize("dim", [84, 42])
def test_warp_matches_cv2_ponglike(dim):
    rng = np.random.default_rng(1)
    for _ in range(40):
        p = np.full((210, 160), 255, np.uint8)
        p[34:194] = 0
        bx, by = rng.integers(0, 157), rng.integers(34, 191)
        p[by:by + 4, bx:bx + 4] = 255
        ly, ry = rng.integers(34, 180, 2)
        p[ly:ly + 15, 16:21] = 255
        p[ry:ry + 15, 139:144] = 255
        p[13:29, 20:140] = rng.integers(0, 256, (16, 120), dtype=np.uint8)  # "text"
        rgb = np.repeat(p[:, :, None], 3, axis=2)
        ref = cv2.resize(cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY), (dim, dim), interpolation=cv2.INTER_AREA)
        assert np.array_equal(warp(rgb, dim), ref)


def test_resize_area_generic_shapes():
    rng = np.random.default_rng(2)
    for (sh, sw, dh, dw) in [(210, 160, 84, 84), (210, 160, 42, 42), (96, 96, 42, 42), (100, 77, 33, 50)]:
        g = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        ref = cv2.resize(g, (dw, dh), interpolation=cv2.INTER_AREA)
        assert np.array_equal(resize_area(g, dw, dh), ref), (sh, sw, dh, dw)


def test_float32_path_restatement_matches_cv2():
    """The float32 observation mode (stock gym) restates cv2's FLOAT paths: RGB2GRAY on float32 is
    fma(B, 0.114f, fma(R, 0.299f, G * 0.587f)) and INTER_AREA on float32 is the uint8 tap arithmetic without the final
    rounding.  csrc/pong_raster_dev.cuh (text_gray_f32, eval_pixel_sum) implements exactly the numpy below."""
    import math
    cv2 = pytest.importorskip("cv2")
    f = np.float32
    cr, cg, cb = f(0.299), f(0.587), f(0.114)

    def fma(a, b, c):
        return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)

    def gray(im):
        return fma(im[..., 2], cb, fma(im[..., 0], cr, im[..., 1] * cg))

    def tab(ssize, dsize):
        scale, out = ssize / dsize, []
        for dx in range(dsize):
            f1 = dx * scale
            f2, cell = f1 + scale, min(scale, ssize - f1)
            s1 = math.ceil(f1)
            s2 = min(math.floor(f2), ssize - 1)
            s1 = min(s1, s2)
            if s1 - f1 > 1e-3:
                out.append((dx, s1 - 1, f((s1 - f1) / cell)))
            for s in range(s1, s2):
                out.append((dx, s, f(1.0 / cell)))
            if f2 - s2 > 1e-3:
                out.append((dx, s2, f(min(min(f2 - s2, 1.0), cell) / cell)))
        return out

    def area(S, dim):
        xt, yt = tab(S.shape[1], dim), tab(S.shape[0], dim)
        D, first = np.zeros((dim, dim), np.float32), [True] * dim
        for dy, sy, beta in yt:
            buf = np.zeros(dim, np.float32)
            for dx, sx, alpha in xt:
                buf[dx] = f(buf[dx] + f(S[sy, sx] * alpha))
            D[dy] = f(beta) * buf if first[dy] else (D[dy] + (f(beta) * buf).astype(np.float32)).astype(np.float32)
            first[dy] = False
        return D
    rng = np.random.default_rng(1)
    for w in list(range(1, 40)) + [160]:           # vector body and scalar tail of cv2's loop agree
        im = rng.integers(0, 256, (5, w, 3)).astype(np.float32)
        assert np.array_equal(gray(im), cv2.cvtColor(im, cv2.COLOR_RGB2GRAY)), w
    assert np.array_equal(cv2.cvtColor(np.full((2, 8, 3), 255, np.float32), cv2.COLOR_RGB2GRAY), np.full((2, 8), 255, np.float32))
    for dim in (84, 42):
        im = np.zeros((210, 160, 3), np.float32)
        im[:34] = 255
        im[194:] = 255
        im[100:104, 50:54] = 255
        im[60:75, 16:21] = 255
        im[8:28, 20:140] = rng.integers(0, 256, (20, 120, 1)).astype(np.float32)      # antialiased text stand-in, R = G = B
        for img in (im, rng.integers(0, 256, (210, 160, 3)).astype(np.float32)):
            want = cv2.resize(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), (dim, dim), interpolation=cv2.INTER_AREA)
            assert want.dtype == np.float32 and np.array_equal(area(gray(img), dim), want), dim
