"""Device-resident throughput of the Pong path in its other observation modes:
resized_dim 84 / 42, frame_stack 4 / None, cPong-v0 / cPongDouble-v0.  One JSON line per variant."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from competitive_rl_b200 import _native, make_envs

lib = _native.load()
N, STEPS = 65536, 300
for env_id, dim, fs in [("cPongDouble-v0", 84, 4), ("cPongDouble-v0", 42, 4), ("cPongDouble-v0", 84, None), ("cPongDouble-v0", 42, None),
                        ("cPong-v0", 84, 4), ("cPong-v0", 42, 4)]:
    envs = make_envs(env_id, seed=1000, log_dir=None, num_envs=N, asynchronous=True, resized_dim=dim, frame_stack=fs, n_buffers=1)
    envs.reset()
    A = 2 if "Double" in env_id else 1
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    actions = torch.zeros((N, A), dtype=torch.int32, device="cuda")
    obs = envs._obs
    o0, o1 = obs[0], (obs[1] if A == 2 else None)

    def one(t):
        _native.check(lib.crl_pong_random_actions(P(actions), A * N, 7, t, sp))
        _native.check(lib.crl_pong_step_state(envs._h, P(actions), P(envs._rew), P(envs._done), P(envs._steps), P(envs._real), sp))
        _native.check(lib.crl_pong_render_obs(envs._h, P(o0), P(o1) if o1 is not None else None, sp))
    for t in range(20):
        one(t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for t in range(STEPS):
        one(20 + t)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    bytes_step = A * (fs or 1) * dim * dim
    print(json.dumps({"env": env_id, "resized_dim": dim, "frame_stack": fs, "envs": N, "ms_per_step": ms,
                      "env_steps_per_s": N / (ms / 1e3), "bytes_per_env_step": bytes_step,
                      "obs_GBps": bytes_step * N / (ms / 1e3) / 1e9}))
    envs.close()
