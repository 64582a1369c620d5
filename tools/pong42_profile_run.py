import ctypes, sys
sys.path.insert(0, '/root/repo')
import torch
from competitive_rl_b200 import _native, make_envs
lib = _native.load()
N = 65536
envs = make_envs("cPongDouble-v0", seed=1000, log_dir=None, num_envs=N, asynchronous=True, resized_dim=42, frame_stack=4, n_buffers=1)
envs.reset()
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
actions = torch.zeros((N, 2), dtype=torch.int32, device="cuda")
o0, o1 = envs._obs
for t in range(12):
    _native.check(lib.crl_pong_random_actions(P(actions), 2 * N, 7, t, sp))
    _native.check(lib.crl_pong_step_state(envs._h, P(actions), P(envs._rew), P(envs._done), P(envs._steps), P(envs._real), sp))
    _native.check(lib.crl_pong_render_obs(envs._h, P(o0), P(o1), sp))
torch.cuda.synchronize()
