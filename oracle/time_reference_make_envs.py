"""Time the REFERENCE's own CPU path -- make_envs("cPongDouble-v0", num_envs=8, asynchronous=True, resized_dim=84,
frame_stack=None, log_dir=None) -> SubprocVecEnv of the reference's Python envs (BASELINE config 1) --
and the in-process DummyVecEnv, in the build container.  Executed verbatim from /root/reference: make_envs.py,
utils/subproc_vec_env.py, utils/dummy_vec_env.py, utils/atari_wrappers.py, pong/base_pong_env.py, with the real cv2;
gym and pygame are the stand-ins of oracle/ref_shim (neither is installable here), which restate pygame.Rect / draw.rect /
font blit with numpy -- so the renderer's share of the time is the stand-in's, not SDL's (a stock install renders through
SDL's C blitters; SURVEY.md section 6 has the caveat).  The reference cannot travel to the GPU box, so the result is a
committed artefact:  python oracle/time_reference_make_envs.py > profiles/r02_reference_make_envs_cpu.json

The stand-ins are installed at module top level because SubprocVecEnv's forkserver workers re-import this script."""
import json
import os
import platform
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

ref_loader.install()
import types  # noqa: E402

if "competitive_rl.register" not in sys.modules:      # register.py pulls car_racing (Box2D): only Pong is timed here
    reg = types.ModuleType("competitive_rl.register")

    def register_competitive_envs():
        import contextlib
        import io
        from competitive_rl.pong.register import register_pong
        with contextlib.redirect_stdout(io.StringIO()):
            register_pong()
    reg.register_competitive_envs = register_competitive_envs
    sys.modules["competitive_rl.register"] = reg
    import competitive_rl.utils as U
    from competitive_rl.utils.atari_wrappers import make_env_a2c_atari
    from competitive_rl.utils.dummy_vec_env import DummyVecEnv
    from competitive_rl.utils.subproc_vec_env import SubprocVecEnv
    U.DummyVecEnv, U.SubprocVecEnv, U.make_env_a2c_atari = DummyVecEnv, SubprocVecEnv, make_env_a2c_atari
    # make_envs imports TournamentEnvWrapper (-> builtin_policies -> torch networks); it is not on the timed path
    cpe = types.ModuleType("competitive_rl.pong.competitive_pong_env")
    cpe.TournamentEnvWrapper = None
    sys.modules["competitive_rl.pong.competitive_pong_env"] = cpe
    # the real package registers its envs when `competitive_rl` is imported (competitive_rl/__init__.py), which is what
    # a worker process does when it unpickles an env thunk; the stub package has no __init__, so do it here
    register_competitive_envs()


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def run(asynchronous, num_envs, steps, warmup):
    import cv2
    import numpy as np
    import competitive_rl.make_envs as ME
    envs = ME.make_envs("cPongDouble-v0", seed=0, log_dir=None, num_envs=num_envs, asynchronous=asynchronous,
                        resized_dim=84, frame_stack=None)
    envs.reset()
    rng = np.random.default_rng(0)
    for _ in range(warmup):
        envs.step(rng.integers(0, 3, (num_envs, 2)))
    t0 = time.perf_counter()
    for _ in range(steps):
        envs.step(rng.integers(0, 3, (num_envs, 2)))
    dt = time.perf_counter() - t0
    envs.close()
    return {"vec_env": type(envs).__name__, "num_envs": num_envs, "steps": steps, "seconds": dt,
            "env_steps_per_s": num_envs * steps / dt, "cv2_threads": cv2.getNumThreads()}


if __name__ == "__main__":
    cores = len(os.sched_getaffinity(0))
    out = {"what": "reference make_envs('cPongDouble-v0', num_envs=8, resized_dim=84, frame_stack=None) under oracle/ref_shim "
                   "(gym / pygame stand-ins, real cv2 %s), random actions, uint8 Box" % __import__("cv2").__version__,
           "host": {"cpu": cpu_model(), "cores_available": cores, "python": platform.python_version()},
           "subproc": run(True, 8, 1500, 100), "dummy": run(False, 8, 400, 40)}
    out["env_steps_per_s_per_core"] = out["subproc"]["env_steps_per_s"] / min(cores, 8)
    print(json.dumps(out, indent=1))
