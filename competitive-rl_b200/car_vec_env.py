"""Vec-env front end of the CUDA car-racing simulator (cCarRacing-v0 / cCarRacingDouble-v0).

Mirrors what make_envs builds for these ids (competitive_rl/make_envs.py:101-110):
  single: gym.make (TimeLimit 1000) -> FrameStack(n) -> WrapPyTorch        obs (N, n, 96, 96)
  Double: gym.make -> MultipleFrameStack(n) -> FlattenMultiAgentObservation -> WrapPyTorch
          obs (N, 2n, 96, 96) (player 0's stack, then player 1's), reward = player 0's,
          done = any car done (utils/atari_wrappers.py:308-334)
stepped by a Dummy/Subproc vec-env with auto-reset (utils/dummy_vec_env.py:51-63).
Car-car contacts of the two-car env are resolved on the device (csrc/car_contact.cuh).

Buffer lifetime: the tensors returned by reset()/step() are owned by the env and rotate over `n_buffers` sets; what
step t returned is overwritten by step t + n_buffers (the reference returns fresh copies).  Pass `copy=True` to get
fresh tensors every step, or a larger `n_buffers` to keep a short rollout alive."""
import ctypes

import numpy as np
import torch

from . import _native, spaces
from .vec_env import AlreadySteppingError, NotSteppingError, VecEnv, _EnvList


class LazyCarInfos(object):
    """infos[i] -> {"num_steps": ..} (single) or {player: {"num_steps": .., "reward": ..}} (Double),
    plus "terminal_observation" / "TimeLimit.truncated" for finished envs."""

    def __init__(self, env, num_steps, rewards, done, trunc_bits, term):
        self._env, self.num_steps, self.rewards, self.done = env, num_steps, rewards, done
        # gym TimeLimit: the key exists only on the step the limit fired; its value is `not done` (always False with two
        # cars, whose `done` is a dict -- a quirk of the reference stack this reproduces)
        self._trunc_bits = trunc_bits          # decoded on first use: nothing is launched per step for what nobody reads
        self._term, self._host = term, None

    @property
    def time_limit_hit(self):
        return (self._trunc_bits & 2) != 0

    @property
    def truncated(self):
        return (self._trunc_bits & 1) != 0

    def __len__(self):
        return self._env.num_envs

    def terminal_observation(self):
        return self._term

    def __getitem__(self, i):
        if self._host is None:
            self._host = (self.num_steps.cpu().numpy(), self.rewards.cpu().numpy(), self.done.cpu().numpy(),
                          self.truncated.cpu().numpy(), self.time_limit_hit.cpu().numpy())
        steps, rew, done, trunc, hit = self._host
        if self._env.players == 1:
            info = {"num_steps": int(steps[i])}
        else:
            info = {k: {"num_steps": int(steps[i]), "reward": float(rew[i, k])} for k in range(self._env.players)}
        if done[i]:
            t = self._term[i]
            info["terminal_observation"] = t.cpu().numpy() if self._env.return_numpy else t
        if hit[i]:
            info["TimeLimit.truncated"] = bool(trunc[i])
        return info

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class CudaCarVecEnv(VecEnv):
    """actions: float (N, 2) for cCarRacing-v0, (N, 2, 2) for cCarRacingDouble-v0 (steer, gas/brake in [-1, 1]).
    step -> (obs uint8 (N, players*C, 96, 96), rew float32, done bool (N,), infos).

    stack_mode="stack" (default): a fully materialised stack, as the reference returns it.  The env's observation buffers
    form a rotation of C + max(1, n_buffers - 1) sets registered with the library, which writes every new frame straight
    into the C buffers it will appear in, so no frame is copied from step to step.  "stack-shift" is the plain C-ABI mode
    (any buffer per call; internal frame ring and a stack-shift kernel).
    stack_mode="ring" (opt-in): the observation is a strided VIEW of an (N, players, 2C, 96, 96) double-write ring -- shape
    (N, C, 96, 96) for one car, (N, 2, C, 96, 96) for two (player axis kept: the two players' windows cannot be one
    uniformly strided channel axis; `.flatten(1, 2)` materialises the reference's (N, 2C, 96, 96) layout).  A step then
    writes each new frame twice and copies nothing; the view of step t is valid until step t + 1."""

    def __init__(self, env_id="cCarRacing-v0", num_envs=1, frame_stack=4, action_repeat=None, seed=0,
                 asynchronous=False, device=None, max_episode_steps=1000, first_env=0, track_draws=None, birth=None,
                 glyphs="default", return_numpy=False, n_buffers=2, copy=False, done_mode="any", stack_mode="stack"):
        if env_id not in ("cCarRacing-v0", "cCarRacingDouble-v0"):
            raise ValueError("unsupported env id %r" % (env_id,))
        if not torch.cuda.is_available():
            raise RuntimeError("CudaCarVecEnv needs a CUDA device: this simulator has no CPU path")
        ext = _native.ext()
        self._lib = _native.load()          # the plain C ABI, for callers that drive it directly (bench.py, tests)
        self.env_id, self.players = env_id, 2 if env_id == "cCarRacingDouble-v0" else 1
        self.c = int(frame_stack) if frame_stack else 1
        self.asynchronous, self.return_numpy, self.copy = bool(asynchronous), bool(return_numpy), bool(copy)
        if stack_mode not in ("stack", "ring", "stack-shift"):
            raise ValueError("stack_mode must be 'stack', 'ring' or 'stack-shift'")
        self.stack_mode, self.ring = stack_mode, stack_mode == "ring"
        if done_mode not in ("any", "car0"):
            raise ValueError("done_mode must be 'any' (make_envs) or 'car0' (make_competitive_car_racing)")
        self.max_episode_steps = int(max_episode_steps or 0)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        n = int(num_envs)
        ch = self.players * self.c
        obs_space = spaces.Box(0, 255, (ch, 96, 96), dtype=np.uint8)
        act_space = spaces.Box(-1, 1, (2,) if self.players == 1 else (self.players, 2), dtype=np.float32)
        VecEnv.__init__(self, n, obs_space, act_space)
        self._impl = ext.Car(n, self.players, int(frame_stack or 0), int(action_repeat or 0), self.max_episode_steps,
                             int(self.device.index), 1 if done_mode == "car0" else 0, 1 if self.ring else 0,
                             int(seed) & (2 ** 64 - 1), int(first_env))
        self._h = ctypes.c_void_p(self._impl.raw_handle())
        if isinstance(glyphs, str) and glyphs == "default":
            g = np.load(_native.DEFAULT_CAR_GLYPHS)
            glyphs = np.concatenate([g["bitmaps"].reshape(-1), g["advance"].reshape(-1)]).astype(np.uint8)
        if glyphs is not None:
            self._impl.load_glyphs(torch.from_numpy(np.ascontiguousarray(glyphs, np.uint8)))
        if track_draws is not None:
            self.inject_tracks(track_draws, birth)
        dev = self.device
        self._ring = torch.empty((n, self.players, 2 * self.c, 96, 96), dtype=torch.uint8, device=dev) if self.ring else None
        # "stack": the observation buffers form a rotation the library knows (crl_car_set_obs_rotation), and every new
        # frame is written straight into the C buffers it will appear in -- nothing is copied from step to step.  What a
        # call returned stays intact for n_buffers - 1 further calls (at least one), as with the other modes.
        # "stack-shift": caller-chosen buffer per call, internal frame ring + stack-shift kernel (the plain C-ABI mode).
        self._rotation = stack_mode == "stack" and self.c >= 2
        n_sets = self.c + max(1, int(n_buffers) - 1) if self._rotation else max(1, int(n_buffers))
        self._sets = []
        for _ in range(n_sets):
            self._sets.append(dict(
                obs=self._ring if self.ring else torch.empty((n, ch, 96, 96), dtype=torch.uint8, device=dev),
                term=torch.zeros((n, ch, 96, 96), dtype=torch.uint8, device=dev),
                rew=torch.zeros((n, self.players), dtype=torch.float32, device=dev),
                done=torch.zeros((n,), dtype=torch.bool, device=dev),
                trunc=torch.zeros((n,), dtype=torch.uint8, device=dev),
                steps=torch.zeros((n,), dtype=torch.int32, device=dev)))
        self._cur = 0
        if self._rotation:
            self._impl.set_obs_rotation([b["obs"] for b in self._sets])
        self._actions = torch.zeros((n, self.players, 2), dtype=torch.float32, device=dev)
        self._waiting, self.closed = False, False
        self.envs = _EnvList(self)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def bytes_per_env_step(self):
        """observation bytes written per env-step (36 864 / 73 728 with frame_stack 4; the ring
        writes each new frame twice whatever the stack depth)"""
        return self.players * (2 if self.ring else self.c) * 96 * 96

    @property
    def _store(self):
        """the buffer the rasteriser writes (the observation itself, or the ring)"""
        return [self._sets[self._cur]["obs"]]

    @staticmethod
    def _ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def inject_tracks(self, draws, birth=None):
        """Validation mode: draws (N, K, 24) = np_random.uniform values per _create_track attempt; birth
        (N, Kb, players) = shuffled birth_place_indices per reset."""
        d = np.ascontiguousarray(draws, np.float64)
        assert d.ndim == 3 and d.shape[0] == self.num_envs and d.shape[2] == 24
        b = None if birth is None else torch.from_numpy(np.ascontiguousarray(birth, np.int32))
        self._impl.inject_tracks(torch.from_numpy(d), b)

    def next_set(self):
        """For callers that drive the C ABI themselves (bench.py, tools): the buffer set the NEXT call must be given
        (stack_mode "stack": the next observation buffer of the registered rotation)."""
        self._cur = (self._cur + 1) % len(self._sets)
        return self._sets[self._cur]

    def _obs_of(self, b):
        """what the caller sees of buffer set b: the observation tensor, or the ring's current window"""
        if not self.ring:
            return b["obs"]
        k = self._impl.ring_phase()
        v = self._ring[:, :, k + 1:k + 1 + self.c]
        return v[:, 0] if self.players == 1 else v

    def _out(self, t):
        if self.return_numpy:
            return t.cpu().numpy()
        return t.clone() if self.copy else t

    def load_tracks(self, tracks):
        """CarRacing.reset(use_local_track=...) (car_racing_multi_players.py:376-381) for the whole vec-env: `tracks` is a
        list of recorded tracks, each a JSON path or an (n, 4) [alpha, beta, x, y] / (n, 3) [beta, x, y] array; from the
        next reset on env i replays track i % len(tracks) at every reset.  An empty list returns to generated tracks."""
        arrs = []
        for t in tracks:
            if isinstance(t, str):
                import json
                with open(t, "r", encoding="utf-8") as f:
                    t = json.load(f)
            t = np.asarray(t, np.float64)
            arrs.append(t[:, 1:4] if t.shape[1] == 4 else t)
        pts = np.zeros((max(len(arrs), 1), 512, 3), np.float64)
        counts = np.zeros((max(len(arrs), 1),), np.int32)
        for k, t in enumerate(arrs):
            assert 9 <= len(t) <= 512, "a track has 9..512 points"
            pts[k, :len(t)] = t
            counts[k] = len(t)
        self._impl.load_tracks(torch.from_numpy(pts), torch.from_numpy(counts), len(arrs))

    def record_track(self, env, path=None):
        """CarRacing.reset(record_track_to=...) (:447-451): the current track of `env` in the reference's JSON format
        [[alpha, beta, x, y], ...].  alpha (the polar angle the generator was at when it emitted the point, unused by
        every consumer of the file) is recomputed as atan2 of the previous point.  Returns the list; writes `path` if given."""
        t = self.get_track(env)
        prev = np.roll(t, 1, axis=0)
        alpha = np.mod(np.arctan2(prev[:, 2], prev[:, 1]), 2 * np.pi)
        rows = [[float(alpha[i]), float(t[i, 0]), float(t[i, 1]), float(t[i, 2])] for i in range(len(t))]
        if path is not None:
            import json
            with open(path, "w", encoding="utf-8") as f:
                json.dump(rows, f)
        return rows

    def reset(self):
        self._cur = (self._cur + 1) % len(self._sets)
        b = self._sets[self._cur]
        self._impl.reset(b["obs"])
        self._waiting = False
        return self._out(self._obs_of(b))

    def step_async(self, actions):
        if self._waiting:
            raise AlreadySteppingError()
        if isinstance(actions, torch.Tensor) and actions.dtype == torch.float32 and actions.device == self.device \
                and actions.is_contiguous() and actions.numel() == self._actions.numel():
            a = actions                      # consumed in place
        else:
            a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions, np.float32))
            self._actions.copy_(a.reshape(self._actions.shape), non_blocking=True)
            a = self._actions
        self._cur = (self._cur + 1) % len(self._sets)
        b = self._sets[self._cur]
        self._impl.step(a, b["obs"], b["rew"], b["done"], b["steps"], b["trunc"], b["term"])
        self._waiting = True

    def step_wait(self):
        if not self._waiting:
            raise NotSteppingError()
        self._waiting = False
        b = self._sets[self._cur]
        obs = self._obs_of(b)
        if self.copy:
            b = {k: v.clone() for k, v in b.items() if k != "obs"}
        done = b["done"]
        rew = b["rew"][:, 0]          # FlattenMultiAgentObservation returns r[0]; single: the scalar reward
        infos = LazyCarInfos(self, b["steps"], b["rew"], b["done"], b["trunc"], b["term"])
        if not self.asynchronous:     # DummyVecEnv buffers: (N, 1)
            rew, done = rew[:, None], done[:, None]
        if self.return_numpy:
            rew, done, infos = rew.cpu().numpy(), done.cpu().numpy(), list(infos)
        return self._out(obs), rew, done, infos

    def seed(self, seed=None):
        """VecEnv.seed: env i is seeded with seed + i (dummy_vec_env.py:65-69) and CarRacing.seed returns [seed]
        (car_racing_multi_players.py:248-250).  Here one key re-seeds the whole batch: env i draws its tracks and
        birth places from the Philox stream of (seed, first_env + i)."""
        if seed is not None:
            self._impl.seed(int(seed) & (2 ** 64 - 1))
        return [[None if seed is None else seed + i] for i in range(self.num_envs)]

    def set_elapsed(self, elapsed):
        """Pre-age the envs: TimeLimit._elapsed_steps per env (int (N,)); spreads the truncations of a synchronously
        reset batch over the steps, like a long-running rollout."""
        t = torch.as_tensor(np.asarray(elapsed)).to(self.device, torch.int32).contiguous()
        assert tuple(t.shape) == (self.num_envs,)
        self._impl.set_elapsed(t)
        torch.cuda.current_stream(self.device).synchronize()

    def get_state(self):
        s = torch.empty((self.num_envs * self.players, 24), dtype=torch.float64, device=self.device)
        self._impl.get_state(s)
        return s.reshape(self.num_envs, self.players, 24)

    def set_state(self, state):
        """Debug / tests: put every car into `state` (N, players, 24) as get_state() lays it out (see crl_car_set_state)."""
        s = torch.as_tensor(state, dtype=torch.float64).to(self.device).reshape(self.num_envs * self.players, 24).contiguous()
        self._impl.set_state(s)
        torch.cuda.current_stream(self.device).synchronize()

    def render_state(self):
        """Debug / tests: render the CURRENT state (e.g. after set_state) as one more frame of the stack; returns the
        newest frame of every player, uint8 (N, players, 96, 96)."""
        if self._rotation:
            self._cur = (self._cur + 1) % len(self._sets)
        b = self._sets[self._cur]
        self._impl.render_state(b["obs"])
        return self._obs_of(b).reshape(self.num_envs, self.players, self.c, 96, 96)[:, :, -1]

    def get_track(self, env):
        return self._impl.get_track(int(env)).numpy()

    def episode_stats(self):
        raw = self._impl.stats()
        ep = int(raw[0])
        return {"episodes": ep, "mean_length": raw[1] / ep if ep else 0.0, "mean_tiles": raw[2] / ep if ep else 0.0,
                "resets_without_pregenerated_track": int(raw[3]), "cut_car_polygons": int(raw[4]),
                "pregen_launches": int(raw[5])}

    def get_contacts(self):
        """(int32 [num_envs] touching car-car fixture pairs after the last step, contacts dropped so far)."""
        counts, over = self._impl.contacts()
        return counts.numpy(), int(over)

    def check(self):
        self._impl.check()

    def close(self):
        if not self.closed and self._impl is not None:
            torch.cuda.synchronize(self.device)
            self._impl.close()
            self._h = None
        self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def get_images(self, indices=None, **kwargs):
        obs = self._obs_of(self._sets[self._cur]).reshape(self.num_envs, self.players, self.c, 96, 96)
        return [np.repeat(obs[i, 0, self.c - 1].cpu().numpy()[:, :, None], 3, axis=2) for i in self._get_indices(indices)]

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.envs[i], attr_name) for i in self._get_indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        for i in self._get_indices(indices):
            setattr(self.envs[i], attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return [getattr(self.envs[i], method_name)(*method_args, **method_kwargs) for i in self._get_indices(indices)]
