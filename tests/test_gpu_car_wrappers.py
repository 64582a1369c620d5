"""The car env stacks as the reference's make_envs / make_competitive_car_racing build them, against fixtures recorded
from the reference's OWN wrappers (oracle/gen_golden_car_wrappers.py: gym.make [TimeLimit] -> FrameStack | MultipleFrameStack
+ FlattenMultiAgentObservation -> WrapPyTorch [-> CarRacingWrapper] under the reference's DummyVecEnv, its renderer on the
pygame stand-in).  Pins: channel layout of the stacks, what a reset fills them with, rewards / dones / info dicts incl. the
gym TimeLimit key, auto-reset + terminal_observation, spaces and return shapes.  Tolerances as in test_gpu_car_parity.py
(fp32 solver: state deviation ~1e-3 moves a truncated pixel coordinate now and then)."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def fixture_policy(o):
    """the opponent of the competitive fixture (oracle/gen_golden_car_wrappers.py::policy), batched on the device"""
    f = o[:, -1].double()
    steer = torch.clamp((f[:, :, 48:].mean(dim=(1, 2)) - f[:, :, :48].mean(dim=(1, 2))) / 64.0, -1, 1)
    return torch.stack([steer, torch.full_like(steer, 0.45)], dim=1)


def _mismatch(a, b):
    return float((np.asarray(a) != np.asarray(b)).mean())


@pytest.mark.parametrize("name", ["car_wrappers_single", "car_wrappers_double", "car_wrappers_competitive"])
def test_reference_wrapper_stacks(name):
    from competitive_rl_b200 import make_envs
    from competitive_rl_b200.competitive_car_racing import make_competitive_car_racing
    g = load_golden(name)
    kind, P, limit = str(g["kind"]), int(g["n_players"]), int(g["limit"])
    N, T = g["actions"].shape[1], g["actions"].shape[0]
    kw = dict(track_draws=g["draws"], birth=g["birth"].astype(np.int32), max_episode_steps=limit)
    if kind == "competitive":
        envs = make_competitive_car_racing(fixture_policy, seed=0, num_envs=N, asynchronous=False, frame_stack=4, **kw)
    else:
        envs = make_envs("cCarRacing-v0" if P == 1 else "cCarRacingDouble-v0", num_envs=N, frame_stack=4, log_dir=None,
                         asynchronous=False, **kw)
    assert tuple(envs.observation_space.shape) == tuple(g["obs_space"]) and tuple(envs.action_space.shape) == tuple(g["act_space"])
    o = envs.reset()
    assert tuple(o.shape) == g["reset_obs"].shape and o.dtype == torch.uint8
    mism = [_mismatch(o.cpu().numpy(), g["reset_obs"])]
    term = {tuple(k): i for i, k in enumerate(g["term_idx"].tolist())}
    n_term = 0
    for t in range(T):
        o, r, d, info = envs.step(g["actions"][t].astype(np.float32))
        assert tuple(r.shape) == g["rew"][t].shape and tuple(d.shape) == g["done"][t].shape          # DummyVecEnv: (N, 1)
        assert r.dtype == torch.float32 and d.dtype == torch.bool
        assert np.abs(r.cpu().numpy() - g["rew"][t]).max() <= 1e-4, (name, t)
        assert np.array_equal(d.cpu().numpy(), g["done"][t]), (name, t)
        mism.append(_mismatch(o.cpu().numpy(), g["obs"][t]))
        for i in range(N):
            inf = info[i]
            flat = P == 1 or kind == "competitive"
            assert (inf["num_steps"] if flat else inf[0]["num_steps"]) == int(g["num_steps"][t][i]), (name, t, i)
            want = int(g["truncated"][t][i])
            assert ("TimeLimit.truncated" in inf) == (want >= 0), (name, t, i)
            if want >= 0:
                assert inf["TimeLimit.truncated"] is bool(want), (name, t, i)
            if not flat:
                for k in range(P):
                    assert abs(inf[k]["reward"] - float(g["info_reward"][t][i][k])) <= 1e-4, (name, t, i, k)
            assert ("terminal_observation" in inf) == ((t, i) in term), (name, t, i)
            if (t, i) in term:
                to = inf["terminal_observation"].cpu().numpy()
                assert to.shape == g["term_obs"][term[(t, i)]].shape
                mism.append(_mismatch(to, g["term_obs"][term[(t, i)]]))
                n_term += 1
    assert n_term == len(term) and n_term >= N
    print("%s: pixel mismatch mean %.5f max %.5f over %d observations" % (name, np.mean(mism), np.max(mism), len(mism)))
    assert np.mean(mism) <= 5e-3 and np.max(mism) <= 5e-2, (np.mean(mism), np.max(mism))
    envs.close()


def test_make_competitive_car_racing_signature_and_conventions():
    """make_competitive_car_racing.py:10-12: (opponent_policy, seed=0, num_envs=3, asynchronous=False, frame_stack=4,
    action_repeat=None); DummyVecEnv conventions by default, SubprocVecEnv's with asynchronous=True; done = car 0's."""
    import inspect
    from competitive_rl_b200.competitive_car_racing import make_competitive_car_racing
    sig = inspect.signature(make_competitive_car_racing)
    assert list(sig.parameters)[:6] == ["opponent_policy", "seed", "num_envs", "asynchronous", "frame_stack", "action_repeat"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(seed=0, num_envs=3, asynchronous=False, frame_stack=4, action_repeat=None)
    with pytest.raises(AssertionError):
        make_competitive_car_racing("not callable")
    calls = []

    def opp(o):
        calls.append(tuple(o.shape))
        return torch.zeros((o.shape[0], 2), device=o.device)
    e = make_competitive_car_racing(opp, 100, 5)                      # positional: seed = 100, num_envs = 5
    assert e.num_envs == 5 and tuple(e.action_space.shape) == (2,) and tuple(e.observation_space.shape) == (4, 96, 96)
    o = e.reset()
    assert tuple(o.shape) == (5, 4, 96, 96) and calls == [(5, 4, 96, 96)]
    o, r, d, info = e.step(np.zeros((5, 2), np.float32))
    assert tuple(r.shape) == (5, 1) and tuple(d.shape) == (5, 1) and len(calls) == 2
    assert set(info[0].keys()) == {"num_steps"}
    e.close()
    e = make_competitive_car_racing(opp, num_envs=4, asynchronous=True)
    e.reset()
    o, r, d, info = e.step(torch.zeros((4, 2), device="cuda"))
    assert tuple(r.shape) == (4,) and tuple(d.shape) == (4,)
    e.close()
