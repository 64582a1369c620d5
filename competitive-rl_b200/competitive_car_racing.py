"""make_competitive_car_racing: single-agent view of a two-car race whose car 1 is driven by an opponent
policy (competitive_rl/car_racing/make_competitive_car_racing.py:10-58).

The reference builds `num_envs` thunks `gym.make("cCarRacingDouble-v0") -> MultipleFrameStack -> WrapPyTorch ->
CarRacingWrapper` under a Dummy/Subproc vec-env.  Per env, CarRacingWrapper
  * feeds `{0: action, 1: opponent_action}` to the two-car env, where `opponent_action = opponent_policy(o[1])` was
    computed from the PREVIOUS observation of car 1 (at reset: from the reset observation);
  * returns car 0's view only: `o[0], r[0], d[0], i[0]` -- the env is done when CAR 0 is done (not "any car", which is
    what FlattenMultiAgentObservation does on the make_envs path), or at the TimeLimit, whose plain `True` it expands;
  * `action_space = action_space[0]`.
The thunks are built WITHOUT `action_repeat` (`_make(..., frame_stack=frame_stack)`, :51), so the reference ignores
that argument; so does this function.

Here the wrapper sits on the batched CUDA vec-env; `opponent_policy` is called ONCE per step on car 1's stacked
observations of the whole batch, (N, C, 96, 96) uint8 on the device, and returns (N, 2) actions."""
import numpy as np
import torch

from . import spaces


class CarRacingWrapper(object):
    def __init__(self, envs, opponent_policy):
        assert envs.players == 2, "needs cCarRacingDouble-v0"
        assert callable(opponent_policy)
        self.env, self.opponent_policy = envs, opponent_policy
        self.num_envs = envs.num_envs
        self.observation_space = spaces.Box(0, 255, (envs.c, 96, 96), dtype=np.uint8)
        self.action_space = spaces.Box(-1, 1, (2,), dtype=np.float32)
        self.metadata = envs.metadata
        self.opponent_action = None

    def _split(self, o):
        c = self.env.c
        return o[:, :c], o[:, c:]

    def _opponent(self, opp_obs):
        a = self.opponent_policy(opp_obs)
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a, np.float32))
        return a.to(self.env.device, torch.float32).reshape(self.num_envs, 2)

    def reset(self):
        own, opp = self._split(self.env.reset())
        self.opponent_action = self._opponent(torch.as_tensor(opp))
        return own

    def step(self, action):
        a0 = action if isinstance(action, torch.Tensor) else torch.as_tensor(np.asarray(action, np.float32))
        a0 = a0.to(self.env.device, torch.float32).reshape(self.num_envs, 2)
        o, r, d, info = self.env.step(torch.stack([a0, self.opponent_action], dim=1))
        own, opp = self._split(o)
        # envs that finished were auto-reset: `opp` already holds their reset observation, exactly what the
        # reference's wrapper.reset() feeds the policy (make_competitive_car_racing.py:35-38)
        self.opponent_action = self._opponent(torch.as_tensor(opp))
        return own, r, d, _Car0Infos(info)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def close(self):
        self.env.close()

    def __getattr__(self, name):
        return getattr(self.env, name)


class _Car0Infos(object):
    """i[0] of the two-car info dict: {"num_steps": ...} (+ vec-env keys of finished envs; the terminal observation is
    car 0's stack)."""

    def __init__(self, infos):
        self._infos = infos

    def __len__(self):
        return len(self._infos)

    def __getitem__(self, i):
        full = self._infos[i]
        out = dict(full[0])
        out.pop("reward", None)     # FlattenMultiAgentObservation's addition, not on this path
        if "terminal_observation" in full:
            out["terminal_observation"] = full["terminal_observation"][:self._infos._env.c]
        return out          # "TimeLimit.truncated" sits beside the player keys in the reference's dict: i[0] does not carry it

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def make_competitive_car_racing(opponent_policy, seed=0, num_envs=3, asynchronous=False, frame_stack=4,
                                action_repeat=None, **kwargs):
    """Same signature as the reference (make_competitive_car_racing.py:10-12); extra keyword arguments go to the
    CUDA vec-env (device, return_numpy, track_draws, ...)."""
    assert callable(opponent_policy)
    from .make_envs import make_envs
    del action_repeat               # dropped by the reference's thunks (:51): every env runs with action_repeat=None
    envs = make_envs("cCarRacingDouble-v0", seed=seed, log_dir=None, num_envs=num_envs, asynchronous=asynchronous,
                     frame_stack=frame_stack, action_repeat=None, done_mode="car0", **kwargs)
    return CarRacingWrapper(envs, opponent_policy)
