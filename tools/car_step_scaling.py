import ctypes, os, sys
sys.path.insert(0, '/root/repo')
import torch
from competitive_rl_b200 import _native, make_envs
lib = _native.load()
for env_id, n in [("cCarRacing-v0", 1024), ("cCarRacing-v0", 4096), ("cCarRacing-v0", 9472), ("cCarRacing-v0", 18944), ("cCarRacing-v0", 32768), ("cCarRacing-v0", 65536), ("cCarRacingDouble-v0", 16384)]:
    envs = make_envs(env_id, num_envs=n, frame_stack=4, log_dir=None, seed=1, n_buffers=1, stack_mode="stack-shift")
    envs.reset()
    P = 2 if "Double" in env_id else 1
    stream = torch.cuda.current_stream(); sp = ctypes.c_void_p(stream.cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    b = envs._sets[0]
    a = torch.zeros((n, P, 2), device="cuda"); a[..., 1] = 0.5
    if P == 2:
        a[:, 0, 0] = 0.0   # straight ahead: cars spawn side by side and stay apart
    ts = []
    for t in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _native.check(lib.crl_car_step_state(envs._h, ptr(a), ptr(b["rew"]), ptr(b["done"]), ptr(b["steps"]), ptr(b["trunc"]), sp))
        e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    cnt = envs.get_contacts()[0] if P == 2 else None
    print(env_id, n, "step kernel ms: %.3f" % (sum(ts[20:]) / 20), "touching frac", None if cnt is None else float((cnt > 0).mean()))
    envs.close()
