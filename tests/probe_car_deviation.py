"""Exploratory (not collected by pytest; run by hand on a GPU box): deviation of the CUDA car path from the C oracle
(states, rewards, pixels).  Lives under tests/ because it uses the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import car_oracle as C
from competitive_rl_b200 import make_envs, _native

N, T, P = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.RandomState(3)
draws = np.zeros((N, 4, 24)); tracks = []
for e in range(N):
    tr, bd, d = C.make_track(rng)
    draws[e, :] = d   # every attempt slot holds a succeeding draw set
    tracks.append((tr, bd))
birth = np.tile(np.arange(P)[None, None], (N, 4, 1)).astype(np.int32)
env_id = "cCarRacing-v0" if P == 1 else "cCarRacingDouble-v0"
envs = make_envs(env_id, num_envs=N, frame_stack=4, log_dir=None, track_draws=draws, birth=birth)
glyphs = C.load_glyphs(_native.DEFAULT_CAR_GLYPHS)
orcs = [C.CarOracleEnv(P, 1, glyphs) for _ in range(N)]
obs = envs.reset()
o_or = [o.reset(*tracks[e], list(range(P))) for e, o in enumerate(orcs)]
tg = envs.get_track(0)
print("track diff", np.abs(tg - tracks[0][0][:, 1:]).max(), tg.shape)
og = obs.cpu().numpy()
mm = np.mean([[ (og[e, p*4+3] != o_or[e][p]).mean() for p in range(P)] for e in range(N)])
print("reset obs mismatch frac", mm, "stack equal", np.array_equal(og[:,0], og[:,3]))
arng = np.random.default_rng(0)
steer = np.zeros((N, P)); maxdev = 0; rew_g = np.zeros((N,P)); rew_o = np.zeros((N,P)); mms = []
t0 = time.time()
for t in range(T):
    if t % 20 == 0: steer = arng.uniform(-0.4, 0.4, (N, P))
    gas = np.where((t // 50) % 3 == 2, -0.5, 0.7)
    a = np.stack([steer, np.full((N, P), gas)], axis=-1).astype(np.float32)
    obs, r, d, info = envs.step(a if P == 2 else a[:, 0])
    sg = envs.get_state().cpu().numpy()
    og = obs.cpu().numpy()
    rg = info.rewards.cpu().numpy()
    for e in range(N):
        oo, ro, do, ns = orcs[e].step(a[e].astype(np.float64))
        so = orcs[e].get_state()
        dev = np.abs(sg[e, :, :6] - so[:, :6]).max()
        maxdev = max(maxdev, dev)
        rew_g[e] += rg[e]; rew_o[e] += ro
        if t % 25 == 0:
            mms.append(np.mean([(og[e, p*4+3] != oo[p]).mean() for p in range(P)]))
    if t % 50 == 0:
        print(t, "max state dev so far", maxdev, "tiles g/o", sg[0,0,23], orcs[0].get_state()[0,23], "done", d.cpu().numpy().sum())
print("time", time.time() - t0)
print("max hull state deviation", maxdev)
print("return diff max", np.abs(rew_g - rew_o).max(), "returns", rew_g[:4].ravel(), rew_o[:4].ravel())
print("pixel mismatch mean/max", np.mean(mms), np.max(mms))
envs.check()
import cv2
cv2.imwrite(os.path.join(ROOT, "gpurun_out", "car_gpu.png"), np.concatenate([og[0, 3], oo[0]], 1))
