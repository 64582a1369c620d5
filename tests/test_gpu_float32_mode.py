"""obs_dtype="float32": the observations of a STOCK gym install of the reference (Box without dtype = float32),
against fixtures recorded from the reference's own path run that way (oracle/gen_golden_float32.py, real cv2 float
paths).  Bar: bit-exact float32, including the rounded-integer frames that reset() puts into the stack."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pong_double_84_f32", "pong_single_42_fs4_f32"])
def test_float32_observations_match_reference(name):
    from competitive_rl_b200 import make_envs
    g = load_golden(name)
    env_id = str(g["env_id"])
    double = env_id == "cPongDouble-v0"
    fs = int(g["frame_stack"]) or None
    T, N = g["actions"].shape[:2]
    envs = make_envs(env_id, num_envs=N, resized_dim=int(g["dim"]), frame_stack=fs, log_dir=None, serves=g["serves"],
                     obs_dtype="float32")
    assert envs.observation_space[0].dtype == np.float32 if double else envs.observation_space.dtype == np.float32

    def st(o):
        return np.stack([x.cpu().numpy() for x in o]) if double else o.cpu().numpy()[None]
    o = envs.reset()
    assert st(o).dtype == np.float32 and np.array_equal(st(o), g["reset_obs"])
    term = {tuple(k): i for i, k in enumerate(g["term_idx"].tolist())}
    n_term, n_frac = 0, 0
    for t in range(T):
        o, r, d, info = envs.step(g["actions"][t])
        got = st(o)
        assert np.array_equal(got, g["obs"][t]), (name, t, float(np.abs(got - g["obs"][t]).max()))
        n_frac += int((got != np.rint(got)).sum())
        dd = d.cpu().numpy().reshape(N, -1)[:, 0]
        assert np.array_equal(dd, g["done"][t])
        for i in np.nonzero(dd)[0]:
            to = info[int(i)]["terminal_observation"]
            to = np.stack([x.cpu().numpy() for x in to]) if double else to.cpu().numpy()[None]
            assert np.array_equal(to, g["term_obs"][term[(t, int(i))]]), (name, t, i)
            n_term += 1
    assert n_term == len(term) and n_term >= 1 and n_frac > 1000        # the mode really is unrounded
    envs.close()
