"""Build the scoreboard atlas: the 34 frame rows above the arena (rows 0..33 of the
210x160x3 frame) for every score pair, exactly as the renderer draws them.

The atlas is DATA consumed by the CUDA path (crl_pong_load_atlas) and by the
oracle; it plays the role the font file plays for pygame.  Two back-ends:

  --backend shim     (default) run the reference's own Scoreboard.draw
                     (pong/base_pong_env.py:474-487) under oracle/ref_shim's pygame
                     stand-in, whose glyph coverage comes from PIL/FreeType rendering
                     of FreeSansBold.ttf.  Needs /root/reference (build container).
  --backend pygame   run it under a REAL pygame (1.9.6 per the reference's setup.py:8)
                     wherever one is installed; this is what makes the observation
                     rows 3..13 bit-identical to a stock competitive-rl install.

Output: npz with `strips` uint8 [22][22][34][160][3] (score_left, score_right,
row, col, rgb), `backend`, `note`.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_OUT = os.path.join(ROOT, "competitive-rl_b200", "data", "scoreboard_atlas.npz")


def build(backend):
    if backend == "shim":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_loader
        ref_loader.install()
    import pygame
    if backend == "pygame":
        assert hasattr(pygame, "version"), "a real pygame is required for --backend pygame"
    sys.path.insert(0, os.environ.get("CRL_REFERENCE_ROOT", "/root/reference"))
    from competitive_rl.pong.base_pong_env import Scoreboard, WHITE
    pygame.init()
    surface = pygame.Surface((160, 210))
    board = Scoreboard(20, 8, font_size=20)  # PongGame.__init__, base_pong_env.py:208
    strips = np.full((22, 22, 34, 160, 3), 255, np.uint8)
    for l in range(22):
        for r in range(22):
            surface.fill(WHITE)
            board.draw(surface, l, r)
            img = np.transpose(pygame.surfarray.array3d(surface).astype(np.uint8), (1, 0, 2))
            assert (img[34:] == 255).all(), "scoreboard reaches into the arena rows; atlas rows must grow"
            strips[l, r] = img[:34]
    return strips


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["shim", "pygame"], default="shim")
    ap.add_argument("--out", default=DEFAULT_OUT)
    a = ap.parse_args()
    strips = build(a.backend)
    note = ("glyph coverage from PIL/FreeType FreeSansBold 20px + SDL1.2-style alpha blit (stand-in)"
            if a.backend == "shim" else "rendered by real pygame")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    np.savez_compressed(a.out, strips=strips, backend=a.backend, note=note)
    print("wrote", a.out, strips.shape, "bytes", os.path.getsize(a.out))


if __name__ == "__main__":
    main()
