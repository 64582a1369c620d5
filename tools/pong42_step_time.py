"""Device-timed step of cPongDouble-v0 at 42x42x4 (65 536 envs), a few repetitions: for A/B runs of raster-kernel variants
(build the variant into another .so, copy it over competitive-rl_b200/libcrl_b200.so on the GPU box, run this again)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from competitive_rl_b200 import _native, make_envs

lib = _native.load()
N, STEPS, REPS = 65536, 1000, 4
envs = make_envs("cPongDouble-v0", seed=1000, log_dir=None, num_envs=N, asynchronous=True, resized_dim=42, frame_stack=4, n_buffers=1)
envs.reset()
stream = torch.cuda.current_stream()
sp = ctypes.c_void_p(stream.cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
actions = torch.zeros((N, 2), dtype=torch.int32, device="cuda")
o0, o1 = envs._obs


def one(t):
    _native.check(lib.crl_pong_random_actions(P(actions), 2 * N, 7, t, sp))
    _native.check(lib.crl_pong_step_state(envs._h, P(actions), P(envs._rew), P(envs._done), P(envs._steps), P(envs._real), sp))
    _native.check(lib.crl_pong_render_obs(envs._h, P(o0), P(o1), sp))


t = 0
for _ in range(300):            # past the first points: scores differ from env to env, as in a long rollout
    one(t); t += 1
out = []
for _ in range(REPS):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(STEPS):
        one(t); t += 1
    e1.record(stream)
    torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / STEPS)
print(json.dumps({"tag": os.environ.get("TAG", ""), "ms_per_step": out, "env_steps_per_s": [N / (m / 1e3) for m in out],
                  "checksum": int(o0.to(torch.int64).sum().item()) ^ int(o1.to(torch.int64).sum().item())}))
