"""A short rollout of one configuration, for ncu captures (B200_PROFILING.md):

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/x \
        python tools/profile_run.py --what pong --dim 42 --stack-mode stack --steps 30
    python tools/profile_run.py --what car --double 1 --envs 16384 --steps 30 [--age 1]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", choices=["pong", "car"], default="pong")
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--dim", type=int, default=84)
    ap.add_argument("--stack-mode", default="stack")
    ap.add_argument("--frame-stack", type=int, default=4)
    ap.add_argument("--double", type=int, default=1)
    ap.add_argument("--age", type=int, default=1)
    a = ap.parse_args()
    import numpy as np
    import torch
    from competitive_rl_b200 import _native, make_envs
    lib = _native.load()
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    if a.what == "pong":
        N = a.envs or 65536
        env_id = "cPongDouble-v0" if a.double else "cPong-v0"
        envs = make_envs(env_id, seed=1000, log_dir=None, num_envs=N, asynchronous=True, resized_dim=a.dim,
                         frame_stack=a.frame_stack or None, n_buffers=1, stack_mode=a.stack_mode)
        envs.reset()
        A = 2 if a.double else 1
        actions = torch.zeros((N, A), dtype=torch.int32, device="cuda")
        o0, o1 = envs._store
        for t in range(a.steps):
            _native.check(lib.crl_pong_random_actions(P(actions), A * N, 7, t, sp))
            _native.check(lib.crl_pong_step_state(envs._h, P(actions), P(envs._rew), P(envs._done), P(envs._steps), P(envs._real), sp))
            _native.check(lib.crl_pong_render_obs(envs._h, P(o0), P(o1), sp))
    else:
        N = a.envs or (16384 if a.double else 1024)
        env_id = "cCarRacingDouble-v0" if a.double else "cCarRacing-v0"
        PL = 2 if a.double else 1
        envs = make_envs(env_id, num_envs=N, frame_stack=a.frame_stack, log_dir=None, seed=1, asynchronous=True, n_buffers=2,
                         stack_mode=a.stack_mode)
        envs.reset()
        if a.age:
            envs.set_elapsed(np.random.default_rng(7).integers(0, 1000, N))
        actions = torch.zeros((N, PL, 2), dtype=torch.float32, device="cuda")
        for t in range(a.steps):
            b = envs.next_set()              # stack mode: the next buffer of the rotation the env registered
            _native.check(lib.crl_car_random_actions(P(actions), actions.numel(), 7, t, sp))
            _native.check(lib.crl_car_step(envs._h, P(actions), P(b["obs"]), P(b["rew"]), P(b["done"]), P(b["steps"]), P(b["trunc"]),
                                           P(b["term"]), sp))
    torch.cuda.synchronize()
    envs.close()


if __name__ == "__main__":
    main()
