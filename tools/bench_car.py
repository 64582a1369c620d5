"""Throughput of the CUDA car-racing path (BASELINE configs 4 and 5): env-steps/s with device-side
random actions, CUDA-event timing, per-kernel split.  Prints one JSON line per configuration.

    python tools/bench_car.py [--steps K] [--warmup W] [--gpus N]

--gpus N > 1 re-launches under torch.distributed.run (one rank per GPU); the envs shard by index with no
collective on the step path, times are the max over ranks (BASELINE config 5: 16384 two-car envs per GPU).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(env_id, n, steps, warmup, cpu_envs=0):
    import torch
    import torch.distributed as dist
    from competitive_rl_b200 import _native, make_envs
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    lib = _native.load()
    P = 2 if "Double" in env_id else 1
    envs = make_envs(env_id, num_envs=n, frame_stack=4, log_dir=None, seed=1, asynchronous=True, n_buffers=2,
                     first_env=rank * n, stack_mode=os.environ.get("CRL_STACK_MODE", "stack"))
    envs.reset()
    dev = envs.device
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    class Cur:                # the buffer set of the current call (stack mode: the next buffer of the env's registered rotation)
        b = envs._sets[envs._cur]

        def __getitem__(self, k):
            return self.b[k]
    b = Cur()
    actions = torch.zeros((n, P, 2), dtype=torch.float32, device=dev)
    h = envs._h

    def one_combined(t):      # the call the vec-env makes: crl_car_step (two-car envs: touching cars on a side stream)
        b.b = envs.next_set()
        _native.check(lib.crl_car_random_actions(ptr(actions), actions.numel(), 7, t, sp))
        actions[:, :, 0].mul_(0.3)   # keep cars on the road for a realistic mix (still random)
        _native.check(lib.crl_car_step(h, ptr(actions), ptr(b["obs"]), ptr(b["rew"]), ptr(b["done"]), ptr(b["steps"]), ptr(b["trunc"]), None, sp))

    def one(t, ev=None):      # the two halves separately, for the per-kernel split
        b.b = envs.next_set()
        _native.check(lib.crl_car_random_actions(ptr(actions), actions.numel(), 7, t, sp))
        actions[:, :, 0].mul_(0.3)
        if ev:
            ev[0].record(stream)
        _native.check(lib.crl_car_step_state(h, ptr(actions), ptr(b["rew"]), ptr(b["done"]), ptr(b["steps"]), ptr(b["trunc"]), sp))
        if ev:
            ev[1].record(stream)
        _native.check(lib.crl_car_render_obs(h, ptr(b["obs"]), None, sp))
        if ev:
            ev[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t = 0
    for _ in range(warmup):
        one_combined(t)
        t += 1
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(steps):
        one_combined(t)
        t += 1
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for k in range(steps):
        one(t, evs[k])
        t += 1
    barrier()
    phys = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    rend = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
    if world > 1:
        tt = torch.tensor([ms, phys, rend], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, phys, rend = [float(x) for x in tt.tolist()]
    stats = envs.episode_stats()
    bytes_per_step = P * 4 * 96 * 96
    out = {
        "metric": "%s env-steps/sec at 96x96x%d obs" % (env_id, 4 * P), "value": n * world * steps / (ms / 1e3), "unit": "env-steps/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "scaling": "weak",
        "config": {"workload": "%s, %d envs per GPU x %d GPU(s), frame_stack 4, random actions (steer scaled 0.3)" % (env_id, n, world)},
        "kernel_ms": {"car_step_kernel": phys, "render+autoreset": rend,
                      "note": "the two halves run back to back (crl_car_step_state, crl_car_render_obs); value / ms_per_step time crl_car_step"},
        "roofline": {"bound": "latency (serial 180+60-iteration joint solver per car); HBM shown for reference",
                     "achieved": bytes_per_step * n / (rend / 1e3) / 1e9, "unit": "GB/s", "peak": 6539.2,
                     "frac": bytes_per_step * n / (rend / 1e3) / 1e9 / 6539.2},
        "episodes_finished": stats["episodes"], "mean_tiles_per_episode": stats["mean_tiles"],
    }
    envs.close()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    if a.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        sys.exit(subprocess.call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
                                  "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300)] + sys.argv))
    import torch
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    run("cCarRacing-v0", 1024, a.steps, a.warmup)
    run("cCarRacingDouble-v0", 16384, a.steps, a.warmup)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.destroy_process_group()
