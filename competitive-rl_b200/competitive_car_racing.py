"""make_competitive_car_racing: single-agent view of a two-car race whose car 1 is driven by an opponent
policy (competitive_rl/car_racing/make_competitive_car_racing.py:10-67).  The reference wraps ONE Double
env and calls `opponent_policy(o[1])`; here the wrapper sits on the batched vec-env and the policy maps the
opponent's stacked observation (N, C, 96, 96) to actions (N, 2), all on the device."""
import torch


class CarRacingWrapper(object):
    def __init__(self, envs, opponent_policy):
        assert envs.players == 2, "needs cCarRacingDouble-v0"
        self.env, self.opponent_policy = envs, opponent_policy
        self.num_envs = envs.num_envs
        from . import spaces
        import numpy as np
        self.observation_space = spaces.Box(0, 255, (envs.c, 96, 96), dtype=np.uint8)
        self.action_space = spaces.Box(-1, 1, (2,), dtype=np.float32)
        self._opp_obs = None

    def reset(self):
        o = self.env.reset()
        c = self.env.c
        self._opp_obs = o[:, c:]
        return o[:, :c]

    def step(self, action):
        a0 = torch.as_tensor(action, dtype=torch.float32, device=self.env.device).reshape(self.num_envs, 2)
        a1 = torch.as_tensor(self.opponent_policy(self._opp_obs), dtype=torch.float32,
                             device=self.env.device).reshape(self.num_envs, 2)
        o, r, d, info = self.env.step(torch.stack([a0, a1], dim=1))
        c = self.env.c
        self._opp_obs = o[:, c:]
        return o[:, :c], r, d, info

    def close(self):
        self.env.close()


def make_competitive_car_racing(opponent_policy, num_envs=1, seed=0, frame_stack=4, action_repeat=None, **kwargs):
    from .make_envs import make_envs
    envs = make_envs("cCarRacingDouble-v0", seed=seed, log_dir=None, num_envs=num_envs, asynchronous=True,
                     frame_stack=frame_stack, action_repeat=action_repeat, **kwargs)
    return CarRacingWrapper(envs, opponent_policy)
