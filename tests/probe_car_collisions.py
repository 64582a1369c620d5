"""Exploratory (not collected by pytest; run by hand on a GPU box): cCarRacingDouble car-car collisions, CUDA path vs
the C oracle (states, contact counts).  Lives under tests/ because it uses the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import car_oracle as C
from competitive_rl_b200 import make_envs

SCEN = [(-0.35, 0.35, 0.5, 0.5), (-0.2, 0.3, 0.6, 0.4), (-0.5, 0.0, 0.5, 0.3), (0.0, 0.45, 0.3, 0.6),
        (-0.3, 0.3, 0.8, 0.8), (-0.15, 0.15, 0.4, 0.4), (-0.4, 0.1, 0.7, 0.2), (-0.1, 0.4, 0.2, 0.7)]
N, T = len(SCEN), int(sys.argv[1]) if len(sys.argv) > 1 else 90
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.RandomState(5)
draws = np.zeros((N, 4, 24)); tracks = []
for e in range(N):
    tr, bd, d = C.make_track(rng)
    draws[e, :] = d
    tracks.append((tr, bd))
birth = np.tile(np.arange(2)[None, None], (N, 4, 1)).astype(np.int32)
envs = make_envs("cCarRacingDouble-v0", num_envs=N, frame_stack=4, log_dir=None, track_draws=draws, birth=birth, action_repeat=REP)
orcs = [C.CarOracleEnv(2, REP, None, render=False) for _ in range(N)]
envs.reset()
for e, o in enumerate(orcs):
    o.reset(*tracks[e], [0, 1])
dev_hist = np.zeros((T, N)); cg_hist = np.zeros((T, N), int); co_hist = np.zeros((T, N), int)
for t in range(T):
    a = np.array([[[s[0], s[2]], [s[1], s[3]]] for s in SCEN], np.float32)
    envs.step(a)
    sg = envs.get_state().cpu().numpy()
    cg, over = envs.get_contacts()
    for e in range(N):
        orcs[e].step(a[e].astype(np.float64))
        so = orcs[e].get_state()
        dev_hist[t, e] = np.abs(sg[e, :, :3] - so[:, :3]).max()
        co_hist[t, e] = orcs[e].contacts()[0]
    cg_hist[t] = cg
for e in range(N):
    first_o = np.argmax(co_hist[:, e] > 0) if (co_hist[:, e] > 0).any() else -1
    first_g = np.argmax(cg_hist[:, e] > 0) if (cg_hist[:, e] > 0).any() else -1
    print("scenario", e, "first contact oracle/gpu", first_o, first_g, "contact steps", (co_hist[:, e] > 0).sum(), (cg_hist[:, e] > 0).sum(),
          "count mismatches", (co_hist[:, e] != cg_hist[:, e]).sum(), "max dev: before %.2e, +10 %.2e, +30 %.2e, end %.2e" % (
              dev_hist[:max(first_o, 1), e].max(), dev_hist[:first_o + 10, e].max(), dev_hist[:first_o + 30, e].max(), dev_hist[:, e].max()))
print("overflow", over)
