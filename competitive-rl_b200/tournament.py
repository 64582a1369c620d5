"""TournamentEnvWrapper: single-agent view of a Double vec-env whose agent 1 is played by a built-in opponent
(competitive_rl/pong/competitive_pong_env.py:9-53).  The opponent acts on the device (builtin_policies.py): its
observation, frame stack and action never visit the host.  Available opponents: RANDOM, RULE_BASED and every
network agent whose checkpoint file can be found (the reference ships WEAK and MEDIUM)."""
import numpy as np
import torch

from .builtin_policies import get_builtin_agent_names, get_compute_action_function


class TournamentEnvWrapper(object):
    def __init__(self, env, num_envs, resource_dir=None):
        self.env = env
        self.num_envs = num_envs
        self._resource_dir = resource_dir
        self.agent_names = [n for n in get_builtin_agent_names(resource_dir) if n != "ALPHA_PONG"]
        self.agents = {}                      # built on first use: a network agent allocates its stack and weights
        self.current_agent_name = "RULE_BASED"
        self.current_agent = self._agent("RULE_BASED")
        self.observation_space = env.observation_space[0]
        self.action_space = env.action_space[0]
        self.prev_opponent_obs = None

    def _agent(self, name):
        if name not in self.agents:
            self.agents[name] = get_compute_action_function(name, self.num_envs, self.env.device, self._resource_dir)
        return self.agents[name]

    def get_agent_names(self):
        return self.agent_names

    def reset_opponent(self, agent_name=None):
        if agent_name is None:
            agent_name = self.agent_names[int(np.random.randint(len(self.agent_names)))]
        assert agent_name in self.agent_names, self.agent_names
        self.current_agent_name = agent_name
        self.current_agent = self._agent(agent_name)

    def step(self, action):
        a = torch.as_tensor(np.asarray(action) if not isinstance(action, torch.Tensor) else action)
        a = a.reshape(-1).to(self.env.device, dtype=torch.int32)
        both = torch.stack([a, self.current_agent(self.prev_opponent_obs).reshape(-1)], dim=1)
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
