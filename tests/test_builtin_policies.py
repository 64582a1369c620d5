"""Built-in tournament opponents (competitive-rl_b200/builtin_policies.py): the network definitions against the
reference's own utils/network.py and its shipped checkpoints (CPU, only where /root/reference exists), and the device
policy against a CPU evaluation of the same weights (GPU)."""
import importlib.util
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

REF = "/root/reference"


def _ref_network():
    path = os.path.join(REF, "competitive_rl", "utils", "network.py")
    if not os.path.isfile(path):
        pytest.skip("reference sources not available")
    spec = importlib.util.spec_from_file_location("_ref_network", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)            # the file imports torch only
    return mod


@pytest.mark.parametrize("name", ["weak", "medium"])
def test_light_network_matches_reference_checkpoint(name):
    from competitive_rl_b200.builtin_policies import LightActorCritic
    R = _ref_network()
    ckpt = os.path.join(REF, "resources", "pong", "checkpoint-%s.pkl" % name)
    if not os.path.isfile(ckpt):
        pytest.skip("checkpoint not available")
    state = torch.load(ckpt, map_location="cpu", weights_only=False)["model"]
    ours, ref = LightActorCritic((4, 42, 42), 3), R.LightActorCritic((4, 42, 42), 3)
    ours.load_state_dict(state, strict=True)
    ref.load_state_dict(state, strict=True)
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (64, 4, 42, 42)).astype(np.float32))
    with torch.no_grad():
        lo, vo = ours(x)
        lr, vr = ref(x)
    assert torch.equal(lo, lr) and torch.equal(vo, vr)


def test_full_network_matches_reference_definition():
    from competitive_rl_b200.builtin_policies import ActorCritic
    R = _ref_network()
    torch.manual_seed(0)
    ref = R.ActorCritic((4, 42, 42), 3)
    ours = ActorCritic((4, 42, 42), 3)
    ours.load_state_dict(ref.state_dict(), strict=True)          # same parameter names and shapes
    x = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (16, 4, 42, 42)).astype(np.float32))
    with torch.no_grad():
        assert torch.equal(ours(x)[0], ref(x)[0]) and torch.equal(ours(x)[1], ref(x)[1])


def test_agent_names_follow_available_checkpoints(tmp_path):
    from competitive_rl_b200.builtin_policies import LightActorCritic, get_builtin_agent_names
    assert get_builtin_agent_names(str(tmp_path)) == ["RANDOM", "RULE_BASED"] or "WEAK" in get_builtin_agent_names(str(tmp_path))
    torch.save({"model": LightActorCritic().state_dict()}, str(tmp_path / "checkpoint-weak.pkl"))
    names = get_builtin_agent_names(str(tmp_path))
    assert "WEAK" in names and "RULE_BASED" in names and "RANDOM" in names


@pytest.mark.gpu
def test_device_policy_and_tournament_with_network_opponent(tmp_path):
    from competitive_rl_b200 import make_envs
    from competitive_rl_b200.builtin_policies import ActorCritic, DevicePolicy, LightActorCritic
    torch.manual_seed(3)
    torch.save({"model": LightActorCritic().state_dict()}, str(tmp_path / "checkpoint-weak.pkl"))
    torch.save({"model": ActorCritic().state_dict()}, str(tmp_path / "checkpoint-strong.pkl"))
    N = 256
    t = make_envs("cPongTournament-v0", num_envs=N, resized_dim=42, log_dir=None, seed=4, resource_dir=str(tmp_path))
    assert {"RANDOM", "RULE_BASED", "WEAK", "STRONG"} <= set(t.get_agent_names())
    o = t.reset()
    for name, light in (("WEAK", True), ("STRONG", False)):
        t.reset_opponent(name)
        gpu_pol = t.current_agent
        assert isinstance(gpu_pol, DevicePolicy)
        cpu_pol = DevicePolicy(N, str(tmp_path / ("checkpoint-%s.pkl" % name.lower())), light, "cpu")
        cpu_pol.stack.copy_(gpu_pol.stack.cpu())
        agree = total = 0
        for k in range(25):
            opp_obs = t.prev_opponent_obs.clone()
            expect = cpu_pol.logits(opp_obs.cpu())                       # same stack update, same weights, on the host
            o, r, d, info = t.step(np.random.randint(0, 3, N))
            assert tuple(o.shape) == (N, 1, 42, 42) and tuple(r.shape) == (N, 1) and tuple(d.shape) == (N, 1)
            assert torch.allclose(gpu_pol.stack.cpu(), cpu_pol.stack)     # the opponent's own frame stack
            got = gpu_pol.model(gpu_pol.stack)[0].cpu()
            assert torch.allclose(got, expect, rtol=1e-3, atol=1e-4)
            top2 = expect.topk(2, dim=1).values
            clear = (top2[:, 0] - top2[:, 1]) > 1e-3
            agree += int((got.argmax(1) == expect.argmax(1))[clear].sum())
            total += int(clear.sum())
        assert total > 0 and agree == total
    t.close()
