"""Per-step latency trace of the car path in steady state (pre-aged envs): where the slow steps are and why.

    python tools/car_latency_probe.py [--envs N] [--steps K] [--double 0|1]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--double", type=int, default=1)
    ap.add_argument("--age", type=int, default=1)
    a = ap.parse_args()
    import numpy as np
    import torch
    from competitive_rl_b200 import _native, make_envs
    lib = _native.load()
    env_id = "cCarRacingDouble-v0" if a.double else "cCarRacing-v0"
    P_ = 2 if a.double else 1
    N = a.envs
    envs = make_envs(env_id, num_envs=N, frame_stack=4, log_dir=None, seed=1, asynchronous=True, n_buffers=1, stack_mode="stack-shift")
    envs.reset()
    if a.age:
        envs.set_elapsed(np.random.default_rng(7).integers(0, 1000, N))
    dev = envs.device
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    b = envs._sets[0]
    actions = torch.zeros((N, P_, 2), dtype=torch.float32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record(stream)
    for t in range(a.steps):
        _native.check(lib.crl_car_random_actions(ptr(actions), actions.numel(), 7, t, sp))
        _native.check(lib.crl_car_step(envs._h, ptr(actions), ptr(b["obs"]), ptr(b["rew"]), ptr(b["done"]), ptr(b["steps"]),
                                       ptr(b["trunc"]), ptr(b["term"]), sp))
        ev[t + 1].record(stream)
    torch.cuda.synchronize()
    lat = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)])
    med = float(np.median(lat))
    slow = np.nonzero(lat > 1.5 * med)[0]
    print(json.dumps({"env_id": env_id, "envs": N, "median_ms": med, "mean_ms": float(lat.mean()), "max_ms": float(lat.max()),
                      "first_10_ms": [round(float(x), 3) for x in lat[:10]],
                      "slow_steps(>1.5x median)": [(int(i), round(float(lat[i]), 3)) for i in slow[:40]], "n_slow": int(len(slow)),
                      "stats": envs.episode_stats()}))
    envs.close()


if __name__ == "__main__":
    main()
