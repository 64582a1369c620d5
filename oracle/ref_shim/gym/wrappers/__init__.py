"""gym.wrappers.TimeLimit restated (gym >= 0.12 semantics; gym itself is an un-vendored, unpinned dependency of the
reference and not installed here) -- TEST INFRASTRUCTURE ONLY.  The car-racing registry entries carry
max_episode_steps=1000 (competitive_rl/car_racing/register.py:14,21), so gym.make wraps them in this."""
from ..core import Wrapper


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps=None):
        Wrapper.__init__(self, env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return observation, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)
