"""Short cCarRacingDouble run for ncu captures: 16384 envs, a few steps (tools/bench_car.py is the timed one)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from competitive_rl_b200 import make_envs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
collide = len(sys.argv) > 3 and sys.argv[3] == "collide"      # steer every pair of cars into each other
envs = make_envs("cCarRacingDouble-v0", num_envs=n, frame_stack=4, log_dir=None, seed=1, n_buffers=1)
envs.reset()
gen = torch.Generator(device="cuda").manual_seed(0)
if collide:
    import numpy as np
    s0 = envs.get_state().cpu().numpy()
    d, ang = s0[:, 1, :2] - s0[:, 0, :2], s0[:, 0, 2]
    right = np.cos(ang) * d[:, 0] + np.sin(ang) * d[:, 1]
    toward = torch.as_tensor(np.sign(right) * -0.3, device="cuda", dtype=torch.float32)
    fixed = torch.zeros((n, 2, 2), device="cuda")
    fixed[:, :, 1] = 0.5
    fixed[:, 0, 0], fixed[:, 1, 0] = -toward, toward
for t in range(steps):
    a = torch.rand((n, 2, 2), generator=gen, device="cuda") * 2 - 1
    a[:, :, 0] *= 0.3
    envs.step(fixed if collide else a)
torch.cuda.synchronize()
cnt, over = envs.get_contacts()
print("envs in contact: %.3f, overflow %d" % ((cnt > 0).mean(), over))
envs.check()
