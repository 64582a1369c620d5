"""competitive_rl_b200: B200-native batched simulator for competitive-rl's vectorised
env-stepping path (cPong-v0 / cPongDouble-v0; car racing: see DESIGN.md).

Public surface mirrors the reference package root (competitive_rl/__init__.py:1-6):
make_envs, register_competitive_envs, register_pong, register_car_racing.
"""
from .registry import register_car_racing, register_competitive_envs, register_pong  # noqa: F401
from .make_envs import make_envs  # noqa: F401

__all__ = ["make_envs", "register_competitive_envs", "register_pong", "register_car_racing"]
