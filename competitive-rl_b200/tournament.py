"""TournamentEnvWrapper: single-agent view of a Double vec-env whose agent 1 is played by a
built-in opponent (competitive_rl/pong/competitive_pong_env.py:9-53).  Only the opponents that
need no network are available on the device: RULE_BASED (the env's own action 999,
pong/base_pong_env.py:116-134) and RANDOM."""
import numpy as np
import torch

from .vec_env import CHEAT_CODES


class TournamentEnvWrapper(object):
    def __init__(self, env, num_envs):
        self.env = env
        self.agent_names = ["RANDOM", "RULE_BASED"]
        self.current_agent_name = "RULE_BASED"
        self.observation_space = env.observation_space[0]
        self.action_space = env.action_space[0]
        self.num_envs = num_envs
        self.prev_opponent_obs = None

    def get_agent_names(self):
        return self.agent_names

    def reset_opponent(self, agent_name=None):
        if agent_name is None:
            agent_name = self.agent_names[int(np.random.randint(len(self.agent_names)))]
        assert agent_name in self.agent_names, self.agent_names
        self.current_agent_name = agent_name

    def _opponent_actions(self, n, device):
        if self.current_agent_name == "RULE_BASED":
            return torch.full((n,), CHEAT_CODES, dtype=torch.int32, device=device)
        return torch.randint(0, 3, (n,), dtype=torch.int32, device=device)

    def step(self, action):
        a = torch.as_tensor(np.asarray(action) if not isinstance(action, torch.Tensor) else action)
        a = a.reshape(-1).to(self.env.device, dtype=torch.int32)
        both = torch.stack([a, self._opponent_actions(a.shape[0], a.device)], dim=1)
        obs, rew, done, info = self.env.step(both)
        self.prev_opponent_obs = obs[1]
        if done.ndim == 2:
            done = done[:, 0]
        return obs[0], rew[:, 0].reshape(-1, 1), done.reshape(-1, 1), info

    def reset(self, **kwargs):
        obs = self.env.reset(**kwargs)
        self.prev_opponent_obs = obs[1]
        return obs[0]

    def seed(self, s):
        self.env.seed(s)

    def close(self):
        self.env.close()
