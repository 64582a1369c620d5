"""Step-kernel time of cCarRacingDouble as a function of how many envs have touching cars (CUDA events)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from competitive_rl_b200 import _native, make_envs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lib = _native.load()
for frac in (0.0, 0.03, 0.25, 1.0):
    birth = np.tile(np.arange(2)[None, None], (n, 4, 1)).astype(np.int32)
    envs = make_envs("cCarRacingDouble-v0", num_envs=n, frame_stack=4, log_dir=None, seed=1, n_buffers=1, stack_mode="stack-shift")
    envs.reset()
    s0 = envs.get_state().cpu().numpy()
    # which side is the other car on?  steer towards it for a fraction of the envs, away for the rest
    a = torch.zeros((n, 2, 2), device="cuda")
    a[:, :, 1] = 0.5
    d = s0[:, 1, :2] - s0[:, 0, :2]
    ang = s0[:, 0, 2]
    right = np.cos(ang) * d[:, 0] + np.sin(ang) * d[:, 1]          # car 1 to the right (+) or left (-) of car 0
    toward = torch.as_tensor(np.sign(right) * -0.3, device="cuda", dtype=torch.float32)   # steer action > 0 turns right (steer = -a0)
    hit = torch.rand(n, device="cuda") < frac
    a[:, 0, 0] = torch.where(hit, -toward, toward)
    a[:, 1, 0] = torch.where(hit, toward, -toward)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    b = envs._sets[0]
    times, fracs = [], []
    for t in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _native.check(lib.crl_car_step_state(envs._h, ptr(a), ptr(b["rew"]), ptr(b["done"]), ptr(b["steps"]), ptr(b["trunc"]), sp))
        e1.record(stream)
        torch.cuda.synchronize()
        cnt, over = envs.get_contacts()
        times.append(e0.elapsed_time(e1)); fracs.append(((cnt > 0).mean(), cnt.max(), cnt[cnt > 0].mean() if (cnt > 0).any() else 0))
    for t in (5, 30, 45, 59):
        print("target frac %.2f step %2d: %.3f ms  touching envs %.3f  max contacts %d mean %.2f" % (frac, t, times[t], *fracs[t]))
    envs.close()
