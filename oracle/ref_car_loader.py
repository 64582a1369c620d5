"""Run the reference's OWN car-racing Python (car_dynamics.py, car_racing_multi_players.py) from
/root/reference on top of stand-in Box2D / pygame / gym / matplotlib modules -- TEST
INFRASTRUCTURE ONLY, build container only.

What this pins: the reference's Python-level logic -- Car.step (wheel friction / engine / brake
model), Car.gas/brake/steer, CarRacing.step (rewards, dones, the one-sub-step reward lag),
process_action, FrictionDetector._contact, _create_track -- executed verbatim.
What it does NOT pin: Box2D itself (oracle/ref_shim/Box2D calls oracle/car_oracle.c's solver) and
the observation renderer (pygame polygon fill / rotate are not restated; get_observation is
replaced by a stub below, which is the only modification made to the reference classes).
"""
import os
import sys
import types

import numpy as np

import ref_loader

REFERENCE_ROOT = ref_loader.REFERENCE_ROOT


def install():
    ref_loader.install()   # gym / pygame stand-ins + package stubs
    pkg = sys.modules["competitive_rl"]
    cr = types.ModuleType("competitive_rl.car_racing")
    cr.__path__ = [os.path.join(REFERENCE_ROOT, "competitive_rl", "car_racing")]
    sys.modules["competitive_rl.car_racing"] = cr
    pkg.car_racing = cr


def load_car_racing(render=False):
    """render=False: get_observation is stubbed (fast; logic fixtures).  render=True: the reference's own
    renderer runs on the pygame stand-in (polygon fill / rotate / rect restated from pygame 1.9, see
    oracle/ref_shim/pygame); slow (a 10 000 x 10 000 road map per reset) but it pins the camera, scales,
    colours, paint order and HUD layout of the observation."""
    install()
    import competitive_rl.car_racing.car_racing_multi_players as M
    if not hasattr(M.CarRacing, "_orig_get_observation"):
        M.CarRacing._orig_get_observation = M.CarRacing.get_observation
        M.CarRacing._orig_render_road = M.CarRacing.render_road_for_observation_map
    if render:
        M.CarRacing.get_observation = M.CarRacing._orig_get_observation
        M.CarRacing.render_road_for_observation_map = M.CarRacing._orig_render_road
    else:
        M.CarRacing.get_observation = lambda self, i: np.zeros((96, 96, 1), np.uint8)
        M.CarRacing.render_road_for_observation_map = lambda self, screen: screen
    return M


class RecordingRandom(object):
    """Wraps env.np_random and records every uniform() draw (the track generator's only randomness)."""

    def __init__(self, rng):
        self._rng, self.draws = rng, []

    def uniform(self, a, b):
        v = self._rng.uniform(a, b)
        self.draws.append(float(v))
        return v

    def __getattr__(self, name):
        return getattr(self._rng, name)


def car_state(env, k=0):
    c = env.cars[k]
    h = c.hull
    out = [h.position[0], h.position[1], h.angle, h.linearVelocity[0], h.linearVelocity[1], h.angularVelocity]
    for w in c.wheels:
        out += [w.joint.angle, w.omega, w.gas, len(w.tiles)]
    out += [env.rewards[k], env.tile_visited_count[k]]
    return out
