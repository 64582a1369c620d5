"""Build the HUD glyph atlas for cCarRacing observations: the reward text "%05.0f" is drawn into
every observation with COMIC.TTF at 5 px, non-antialiased (car_racing_multi_players.py:225-229, 669;
pygame_rendering.py:16-18).  Like the Pong scoreboard atlas this is renderer DATA; it is generated
here with PIL/FreeType from the reference's font file (a stand-in for SDL_ttf).

Output: competitive-rl_b200/data/car_hud_glyphs.npz with `bitmaps` uint8 [11][8][4] for
"0123456789-" (row 0 = top of the text line) and `advance` uint8 [11]."""
import os

import numpy as np
from PIL import Image, ImageDraw, ImageFont

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FONT = os.environ.get("CRL_COMIC_TTF", "/root/reference/competitive_rl/car_racing/fonts/COMIC.TTF")
OUT = os.path.join(ROOT, "competitive-rl_b200", "data", "car_hud_glyphs.npz")

if __name__ == "__main__":
    f = ImageFont.truetype(FONT, 5)
    ascent, descent = f.getmetrics()
    assert ascent + descent <= 8
    bitmaps = np.zeros((11, 8, 4), np.uint8)
    advance = np.zeros((11,), np.uint8)
    for i, ch in enumerate("0123456789-"):
        img = Image.new("L", (4, 8), 0)
        d = ImageDraw.Draw(img)
        d.fontmode = "1"
        d.text((0, 0), ch, font=f, fill=255)
        bitmaps[i] = (np.asarray(img) > 0).astype(np.uint8)
        advance[i] = int(round(f.getlength(ch)))
    np.savez_compressed(OUT, bitmaps=bitmaps, advance=advance)
    print("wrote", OUT, advance.tolist())
