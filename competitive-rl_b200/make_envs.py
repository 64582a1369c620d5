"""make_envs: same signature and meaning as competitive_rl/make_envs.py:67-118, returning a
GPU-resident vectorised environment instead of a Dummy/Subproc vec-env of Python envs.

    make_envs(env_id="cPong-v0", seed=0, log_dir="data", num_envs=3, asynchronous=False,
              resized_dim=42, frame_stack=4, action_repeat=None)

Extra keyword-only arguments: device, return_numpy, serves (validation mode), atlas,
first_env (global index of env 0 when the batch is one shard of a multi-GPU job), n_buffers / copy (see below),
stack_mode ("stack" | "ring"), zero_on_done.

BUFFER LIFETIME: unlike the reference's vec-envs, which return fresh numpy copies, the tensors returned by reset() and
step() alias env-owned device buffers that rotate over `n_buffers` (default 2) sets: what step t returned is overwritten
by step t + n_buffers (stack_mode="ring": the observation view by step t + 1).  A rollout that keeps references
(`obs_list.append(obs)`, deferred reads of `infos`) must pass `copy=True` (fresh tensors every step) or a large enough
`n_buffers`.  `infos[i]["terminal_observation"]` is rendered on first access from the episode-end frame specs of the
LAST step: read it before stepping again.
"""
import os
import warnings

from .registry import DEPRECATED_IDS, register_competitive_envs, spec

register_competitive_envs()   # make_envs.py:36 registers at import

__all__ = ["make_envs"]


def _verify_env_id(env_id):
    if env_id in DEPRECATED_IDS:
        new = DEPRECATED_IDS[env_id]
        warnings.warn("Environment id {} is deprecated. Please use the short version {}.".format(env_id, new))
        env_id = new
    return env_id


def make_envs(env_id="cPong-v0", seed=0, log_dir="data", num_envs=3, asynchronous=False, resized_dim=42,
              frame_stack=4, action_repeat=None, **kwargs):
    """
    :param env_id: "cPong-v0", "cPongDouble-v0", "cPongTournament-v0" (deprecated long names accepted)
    :param seed: random seed (serve RNG; env i of the batch uses the stream of global env index i)
    :param log_dir: created if given, otherwise unused (as in the reference, where Monitor is commented out)
    :param num_envs: number of concurrent environments (all stepped by one kernel launch pair)
    :param asynchronous: selects the SubprocVecEnv return conventions (done (N,), Double rew (N, 2))
        instead of DummyVecEnv's (done (N, A), rew (N, A)); like the reference it is forced to False when
        num_envs == 1 (make_envs.py:83)
    :param resized_dim: observation is (C, resized_dim, resized_dim)
    :param frame_stack: frames per observation or None.  For cPongDouble the reference requires None
        (make_envs.py:105-106); here an int stacks per agent with FrameStack's semantics.
    :return: a vectorised environment
    """
    from .vec_env import CudaPongVecEnv
    asynchronous = asynchronous and num_envs > 1
    env_id = _verify_env_id(env_id)
    if env_id == "cPongTournament-v0":
        from .tournament import TournamentEnvWrapper
        resource_dir = kwargs.pop("resource_dir", None)     # where the reference's checkpoint-*.pkl files are (optional)
        envs = make_envs("cPongDouble-v0", seed, log_dir, num_envs, asynchronous, resized_dim, None, **kwargs)
        return TournamentEnvWrapper(envs, num_envs, resource_dir)
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
    if env_id in ("cPong-v0", "cPongDouble-v0"):
        s = spec(env_id)
        kwargs.setdefault("max_num_rounds", s["kwargs"]["max_num_rounds"])   # gym registry kwarg (pong/register.py:13-22)
        return CudaPongVecEnv(env_id, num_envs, resized_dim=resized_dim, frame_stack=frame_stack, seed=seed,
                              asynchronous=asynchronous, **kwargs)
    if env_id in ("cCarRacing-v0", "cCarRacingDouble-v0"):
        from .car_vec_env import CudaCarVecEnv
        s = spec(env_id)
        kwargs.setdefault("max_episode_steps", s.get("max_episode_steps", 1000))   # gym TimeLimit of the registry entry
        return CudaCarVecEnv(env_id, num_envs, frame_stack=frame_stack, action_repeat=action_repeat, seed=seed,
                             asynchronous=asynchronous, **kwargs)
    raise ValueError("unknown environment id %r" % (env_id,))
