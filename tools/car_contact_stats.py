"""cCarRacingDouble under uniform random actions (the steady-state workload of bench.py): how many envs have near /
touching cars, how many contacts, and what the physics pass costs."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from competitive_rl_b200 import _native, make_envs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lib = _native.load()
envs = make_envs("cCarRacingDouble-v0", num_envs=n, frame_stack=4, log_dir=None, seed=1, n_buffers=2)
envs.reset()
envs.set_elapsed(np.random.default_rng(7).integers(0, 1000, n))
stream = torch.cuda.current_stream()
sp = ctypes.c_void_p(stream.cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
actions = torch.zeros((n, 2, 2), dtype=torch.float32, device="cuda")
for t in range(600):
    b = envs.next_set()
    _native.check(lib.crl_car_random_actions(P(actions), actions.numel(), 7, t, sp))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _native.check(lib.crl_car_step(envs._h, P(actions), P(b["obs"]), P(b["rew"]), P(b["done"]), P(b["steps"]), P(b["trunc"]), P(b["term"]), sp))
    e1.record(stream)
    if t % 100 == 99:
        torch.cuda.synchronize()
        cnt, over = envs.get_contacts()
        hist = np.bincount(cnt, minlength=9)[:9]
        print("step %3d: %.3f ms; envs touching %.4f; contacts per touching env mean %.2f max %d; histogram %s" %
              (t, e0.elapsed_time(e1), (cnt > 0).mean(), cnt[cnt > 0].mean() if (cnt > 0).any() else 0, cnt.max(), hist.tolist()))
envs.close()
