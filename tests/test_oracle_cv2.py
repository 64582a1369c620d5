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
