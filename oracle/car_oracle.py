"""ctypes front-end of oracle/car_oracle.c -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcar_oracle.so")
_lib = None
MAX_TRACK = 512
CHECKPOINTS = 12
TRACK_RAD = 900 / 6.0


def build(force=False):
    src = os.path.join(_HERE, "car_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libcar_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp = ctypes.c_void_p
        L.car_oracle_create.restype = vp
        L.car_oracle_create.argtypes = [ctypes.c_int, ctypes.c_int, vp]
        L.car_oracle_destroy.argtypes = [vp]
        L.car_oracle_reset.argtypes = [vp, vp, vp, ctypes.c_int, vp]
        L.car_oracle_step.argtypes = [vp, vp, vp, vp, vp]
        L.car_oracle_obs.restype = ctypes.POINTER(ctypes.c_uint8)
        L.car_oracle_obs.argtypes = [vp, ctypes.c_int]
        L.car_oracle_get_state.argtypes = [vp, vp]
        L.car_oracle_set_lazy_render.argtypes = [vp, ctypes.c_int]
        L.car_oracle_set_state.argtypes = [vp, vp]
        L.car_oracle_env_contacts.argtypes = [vp, vp]
        f = ctypes.c_float
        L.car_oracle_collide_fixtures.argtypes = [ctypes.c_int, f, f, f, ctypes.c_int, f, f, f, vp]
        L.car_oracle_fixture_polygon.argtypes = [ctypes.c_int, vp]
        L.car_oracle_free_collision.argtypes = [vp, vp, ctypes.c_int, vp]
        L.car_oracle_create_track.argtypes = [vp, vp, vp]
        L.car_oracle_create_track.restype = ctypes.c_int
        L.car_oracle_track_border.argtypes = [vp, ctypes.c_int, vp]
        L.car_oracle_polys_touch.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int]
        L.car_oracle_convex_hull.argtypes = [vp, ctypes.c_int, vp]
        L.car_oracle_sizeof_body.restype = ctypes.c_int
        L.car_oracle_sizeof_joint.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def load_glyphs(path):
    g = np.load(path)
    return np.concatenate([g["bitmaps"].reshape(-1), g["advance"].reshape(-1)]).astype(np.uint8)


def draw_track_uniforms(rng):
    """The 24 np_random.uniform draws of one _create_track attempt, in the reference's order
    (car_racing_multi_players.py:268-270): per checkpoint noise ~ U(0, 2pi/12), rad ~ U(R/3, R)."""
    d = np.zeros(2 * CHECKPOINTS)
    for c in range(CHECKPOINTS):
        d[2 * c] = rng.uniform(0, 2 * np.pi * 1 / CHECKPOINTS)
        d[2 * c + 1] = rng.uniform(TRACK_RAD / 3, TRACK_RAD)
    return d


def create_track(draws):
    """-> (track (n,4) float64, border (n,) int32) or None when the attempt fails (the reference retries)."""
    draws = np.ascontiguousarray(draws, np.float64)
    out = np.zeros((MAX_TRACK, 4), np.float64)
    border = np.zeros((MAX_TRACK,), np.int32)
    n = lib().car_oracle_create_track(_p(draws), _p(out), _p(border))
    if n <= 0:
        return None
    return out[:n].copy(), border[:n].copy()


def track_border(track):
    """Kerb flags (:383-397) of a given track (n, 4) [alpha, beta, x, y], e.g. one loaded from the reference's JSON."""
    track = np.ascontiguousarray(track, np.float64)
    border = np.zeros((len(track),), np.int32)
    lib().car_oracle_track_border(_p(track), len(track), _p(border))
    return border


def make_track(rng):
    """Retry like CarRacing.reset (:499-507) until an attempt succeeds; returns (track, border, draws_used)."""
    while True:
        d = draw_track_uniforms(rng)
        t = create_track(d)
        if t is not None:
            return t[0], t[1], d


def collide_fixtures(fa, pose_a, fb, pose_b):
    """b2CollidePolygons + world manifold of two car fixtures (0..3 hull polygons, 4 wheel box) on bodies at
    pose = (x, y, angle).  -> dict(count, type, normal (2,), points (count, 2), ids (count,))."""
    out = np.zeros(10, np.float64)
    n = lib().car_oracle_collide_fixtures(fa, pose_a[0], pose_a[1], pose_a[2], fb, pose_b[0], pose_b[1], pose_b[2], _p(out))
    return {"count": n, "type": int(out[1]), "normal": out[2:4].copy(), "points": out[4:8].reshape(2, 2)[:n].copy(),
            "ids": out[8:10][:n].astype(np.int64)}


def fixture_polygon(f):
    xy = np.zeros((8, 2), np.float32)
    n = lib().car_oracle_fixture_polygon(f, _p(xy))
    return xy[:n].copy()


def free_collision(pose, vel, steps):
    """Two cars in free flight; -> (steps, 9): momentum x, y, angular momentum, contacts, points, hull positions."""
    pose = np.ascontiguousarray(pose, np.float64)
    vel = np.ascontiguousarray(vel, np.float64)
    out = np.zeros((steps, 9), np.float64)
    lib().car_oracle_free_collision(_p(pose), _p(vel), steps, _p(out))
    return out


class CarOracleEnv(object):
    """One cCarRacing env (1 or 2 cars): reset(track, border) -> obs list; step(actions (n_cars, 2))."""

    def __init__(self, n_cars=1, action_repeat=1, glyphs=None, render=True):
        """render=False: reset()/step() return obs=None and frames are drawn only by observe() (the faithful
        per-pixel renderer costs ~0.1 s per frame)."""
        self.n_cars = n_cars
        self.render = bool(render)
        self._glyphs = None if glyphs is None else np.ascontiguousarray(glyphs, np.uint8)
        self._h = lib().car_oracle_create(n_cars, action_repeat, _p(self._glyphs))
        lib().car_oracle_set_lazy_render(self._h, 0 if self.render else 1)

    def close(self):
        if self._h:
            lib().car_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def contacts(self):
        """(touching car-car contacts, manifold points) after the last step."""
        n = ctypes.c_int(0)
        c = lib().car_oracle_env_contacts(self._h, ctypes.byref(n))
        return int(c), int(n.value)

    def observe(self):
        """Render (if lazy) and return the current observation of every player."""
        return [np.ctypeslib.as_array(lib().car_oracle_obs(self._h, k), shape=(96, 96)).copy() for k in range(self.n_cars)]

    def _obs(self):
        if not self.render:
            return None
        return [np.ctypeslib.as_array(lib().car_oracle_obs(self._h, k), shape=(96, 96)).copy() for k in range(self.n_cars)]

    def reset(self, track, border, birth_place=None):
        track = np.ascontiguousarray(track, np.float64)
        border = np.ascontiguousarray(border, np.int32)
        bp = None if birth_place is None else np.ascontiguousarray(birth_place, np.int32)
        lib().car_oracle_reset(self._h, _p(track), _p(border), len(track), _p(bp))
        return self._obs()

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float64).reshape(self.n_cars, 2)
        rew = np.zeros(self.n_cars, np.float64)
        done = np.zeros(self.n_cars, np.int32)
        ns = ctypes.c_int(0)
        lib().car_oracle_step(self._h, _p(a), _p(rew), _p(done), ctypes.byref(ns))
        return self._obs(), rew, done.astype(bool), ns.value

    def set_state(self, state):
        """mirror of crl_car_set_state (tests: renderer cross-checks on arbitrary states)"""
        s = np.ascontiguousarray(state, np.float64).reshape(self.n_cars, 24)
        lib().car_oracle_set_state(self._h, _p(s))

    def get_state(self):
        s = np.zeros((self.n_cars, 24), np.float64)
        lib().car_oracle_get_state(self._h, _p(s))
        return s
