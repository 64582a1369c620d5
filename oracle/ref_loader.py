"""Load the reference's OWN Pong path from /root/reference under stand-in
gym/pygame modules -- TEST INFRASTRUCTURE ONLY.

This only works in the build container (where /root/reference is mounted); it
is used to (1) validate the CPU restatement in oracle/pong_oracle.{c,py} and
(2) generate the committed fixtures under tests/golden/ (oracle/gen_golden.py).
Nothing on the GPU box may import this module's reference path.

Executed verbatim from the reference: pong/base_pong_env.py (game, renderer),
utils/atari_wrappers.py (MaxAndSkip/WarpFrame/ClipReward/FrameStack/WrapPyTorch),
utils/dummy_vec_env.py, utils/subproc_vec_env.py, make_envs.py, with the real
cv2.  Restated by the stand-ins: pygame.Rect/Surface/draw/font/blit, gym spaces
and the underscore-method dispatch (see oracle/ref_shim/*/__init__.py).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CRL_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "competitive_rl", "pong"))


def install():
    """Put the stand-ins on sys.path and stub the package roots whose
    __init__ would import Box2D/matplotlib (competitive_rl/__init__.py:1-6,
    car_racing/__init__.py)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if "competitive_rl" not in sys.modules:
        pkg = types.ModuleType("competitive_rl")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "competitive_rl")]
        sys.modules["competitive_rl"] = pkg
        cr = types.ModuleType("competitive_rl.car_racing")
        cr.make_car_racing = cr.make_car_racing_double = None
        cr.register_car_racing = lambda: None
        sys.modules["competitive_rl.car_racing"] = cr
        pong = types.ModuleType("competitive_rl.pong")
        pong.__path__ = [os.path.join(REFERENCE_ROOT, "competitive_rl", "pong")]
        sys.modules["competitive_rl.pong"] = pong
        utils = types.ModuleType("competitive_rl.utils")
        utils.__path__ = [os.path.join(REFERENCE_ROOT, "competitive_rl", "utils")]
        sys.modules["competitive_rl.utils"] = utils


class ServeInjector(object):
    """Replacement for the module attribute `random` of
    competitive_rl.pong.base_pong_env (stdlib random: base_pong_env.py:1,
    Ball.reset :314-320 draws uniform, choice, choice).

    serves[env][k] = (vx, vy) is the k-th serve env `env` consumes since its
    construction.  `current` selects the env whose stream is being consumed; the
    per-env wrapper below sets it around every call into that env.
    """

    def __init__(self, serves):
        self.serves = serves           # array-like (N, K, 2) float64
        self.count = [0] * len(serves)
        self.current = 0
        self._cur = None
        self._k = 0

    def uniform(self, a, b):
        e = self.current
        self._cur = self.serves[e][self.count[e]]
        self.count[e] += 1
        self._k = 0
        v = abs(float(self._cur[1]))
        assert a <= v <= b, (a, v, b)
        return v

    def choice(self, seq):
        v = float(self._cur[self._k])
        self._k += 1
        return seq[0] if v < 0 else seq[1]


class _Scoped(object):
    """Delegating proxy that tags which env's serve stream is live. Not a gym
    wrapper on purpose: it adds no behaviour to the wrapped reference stack."""

    def __init__(self, injector, rank, thunk):
        object.__setattr__(self, "_inj", injector)
        object.__setattr__(self, "_rank", rank)
        injector.current = rank
        object.__setattr__(self, "_env", thunk())

    def __getattr__(self, name):
        attr = getattr(self._env, name)
        if callable(attr) and name in ("step", "reset"):
            def call(*a, **k):
                self._inj.current = self._rank
                return attr(*a, **k)
            return call
        return attr


def make_reference_vec_env(env_id, num_envs, serves=None, resized_dim=84, frame_stack=None, seed=0):
    """Reference DummyVecEnv over the reference wrapper stack; equivalent to
    make_envs(env_id, seed, None, num_envs, False, resized_dim, frame_stack)
    (make_envs.py:67-118) but with per-env serve injection when `serves` is given.
    Returns (vec_env, injector_or_None)."""
    install()
    import competitive_rl.pong.base_pong_env as B
    from competitive_rl.pong.register import register_pong
    from competitive_rl.utils.atari_wrappers import make_env_a2c_atari
    from competitive_rl.utils.dummy_vec_env import DummyVecEnv
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        register_pong()
    if env_id == "cPongDouble-v0":
        assert frame_stack is None  # make_envs.py:105-106
    thunks = [make_env_a2c_atari(env_id, seed, i, None, resized_dim, frame_stack) for i in range(num_envs)]
    inj = None
    if serves is not None:
        inj = ServeInjector(serves)
        B.random = inj
        thunks = [(lambda i=i, t=t: _Scoped(inj, i, t)) for i, t in enumerate(thunks)]
    else:
        import random as _random
        B.random = _random
    return DummyVecEnv(thunks), inj


def game_of(env):
    """Reach the PongGame under the wrapper stack of one reference env."""
    e = env._env if isinstance(env, _Scoped) else env
    return e.unwrapped._game


def game_state(env):
    """(ball_x, ball_y, vx, vy, left_y, right_y, score_l, score_r, rounds, steps)."""
    g = game_of(env)
    b = g._ball
    return (b._rect.x, b._rect.y, b._speed_x, b._speed_y, g._left_bat._rect.y, g._right_bat._rect.y,
            g._score_left, g._score_right, g._num_rounds, g._num_steps)
