from .. import error

_registry = {}


class EnvSpec(object):
    def __init__(self, id, entry_point=None, kwargs=None, max_episode_steps=None, **_):  # noqa: A002
        self.id = id
        self.entry_point = entry_point
        self.kwargs = dict(kwargs or {})
        self.max_episode_steps = max_episode_steps


def register(id, **kwargs):  # noqa: A002
    if id in _registry:
        raise error.Error("Cannot re-register id: {}".format(id))
    _registry[id] = EnvSpec(id, **kwargs)


def spec(id):  # noqa: A002
    if id not in _registry:
        raise error.UnregisteredEnv("No registered env with id: {}".format(id))
    return _registry[id]


def make(id, **kwargs):  # noqa: A002
    s = spec(id)
    kw = dict(s.kwargs)
    kw.update(kwargs)
    cls = s.entry_point
    if isinstance(cls, str):
        import importlib
        mod, name = cls.split(":")
        cls = getattr(importlib.import_module(mod), name)
    env = cls(**kw)
    env.spec = s
    if s.max_episode_steps is not None:      # gym.envs.registration.EnvSpec.make wraps the env in TimeLimit
        from ..wrappers import TimeLimit
        env = TimeLimit(env, max_episode_steps=s.max_episode_steps)
    return env
