"""Environment-id registry: the role gym's registry plays for the reference
(competitive_rl/register.py:5-7, pong/register.py:8-27, car_racing/register.py:8-26)."""

_REGISTRY = {}

# old spellings accepted by the reference's (commented-out) _verify_env_id, make_envs.py:50-64
DEPRECATED_IDS = {
    "CompetitivePongTournament-v0": "cPongTournament-v0",
    "CompetitivePongDouble-v0": "cPongDouble-v0",
    "CompetitivePong-v0": "cPong-v0",
}


def register(env_id, **spec):
    if env_id in _REGISTRY:
        return False
    _REGISTRY[env_id] = dict(spec)
    return True


def spec(env_id):
    return _REGISTRY[env_id]


def registered_ids():
    return sorted(_REGISTRY)


def register_pong():
    """pong/register.py:8-27: cPong-v0 and cPongDouble-v0 with max_num_rounds=21."""
    a = register("cPong-v0", kind="pong", n_agents=1, kwargs=dict(max_num_rounds=21))
    b = register("cPongDouble-v0", kind="pong", n_agents=2, kwargs=dict(max_num_rounds=21))
    if a or b:
        print("Register cPong-v0 and cPongDouble-v0 environments.")


def register_car_racing():
    """car_racing/register.py:8-26: cCarRacing-v0 / cCarRacingDouble-v0, max_episode_steps=1000."""
    register("cCarRacing-v0", kind="car_racing", n_agents=1, max_episode_steps=1000, kwargs=dict(verbose=0))
    register("cCarRacingDouble-v0", kind="car_racing", n_agents=2, max_episode_steps=1000,
             kwargs=dict(verbose=0, num_player=2))


def register_competitive_envs():
    """competitive_rl/register.py:5-7."""
    register_pong()
    register_car_racing()
