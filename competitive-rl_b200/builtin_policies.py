"""Built-in opponents of the tournament env, evaluated on the device.

Replaces competitive_rl/pong/builtin_policies.py:47-91 + utils/policy_serving.py:10-66 for a batched, GPU-resident
vec-env: the opponent's observation never leaves the device, its frame stack is a device tensor and its action
comes back as a device int32 tensor.  RULE_BASED is the env's own action 999 (pong/base_pong_env.py:116-134),
RANDOM a device randint; WEAK / MEDIUM (LightActorCritic, utils/network.py:73-93) and STRONG / ALPHA_PONG
(ActorCritic, :14-56) are torch networks whose weights come from the reference's checkpoint files
(resources/pong/checkpoint-*.pkl, a dict with the state dict under "model").  Those files are the reference's data,
not part of this repository: point COMPETITIVE_RL_RESOURCES (or the resource_dir argument) at the directory that
holds them; agents whose checkpoint is missing are simply not offered (the reference itself ships only weak and
medium).  The convolutions / linear layers are plain torch (cuDNN / cuBLAS) library calls.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .vec_env import CHEAT_CODES

BUILTIN_AGENT_NAMES = ["RANDOM", "WEAK", "MEDIUM", "STRONG", "RULE_BASED", "ALPHA_PONG"]
_CHECKPOINTS = {"WEAK": ("checkpoint-weak.pkl", True), "MEDIUM": ("checkpoint-medium.pkl", True),
                "STRONG": ("checkpoint-strong.pkl", False), "ALPHA_PONG": ("checkpoint-alphapong.pkl", False)}


class LightActorCritic(nn.Module):
    """4x42x42 -> conv 4x4/2 (16) -> conv 2x2/2 (16) -> 1600 features -> 3 logits, 1 value.  Parameter names are the
    checkpoint's (conv1, conv2, critic_linear, actor_linear)."""

    def __init__(self, input_shape=(4, 42, 42), num_actions=3):
        super().__init__()
        c, h, w = input_shape
        self.conv1 = nn.Conv2d(c, 16, kernel_size=4, stride=2)
        self.conv2 = nn.Conv2d(16, 16, kernel_size=2, stride=2)
        h1, w1 = (h - 4) // 2 + 1, (w - 4) // 2 + 1
        feats = 16 * ((h1 - 2) // 2 + 1) * ((w1 - 2) // 2 + 1)
        self.critic_linear = nn.Linear(feats, 1)
        self.actor_linear = nn.Linear(feats, num_actions)

    def forward(self, x):
        x = F.relu(self.conv1(x / 255.0))
        x = F.relu(self.conv2(x)).flatten(1)
        return self.actor_linear(x), self.critic_linear(x)


class ActorCritic(nn.Module):
    """4x42x42 -> conv 4x4/2 (16) -> conv 4x4/2 pad 2 (32) -> conv 11x11 (256) -> 3 logits, 1 value."""

    def __init__(self, input_shape=(4, 42, 42), num_actions=3):
        super().__init__()
        c, h, w = input_shape
        self.conv1 = nn.Conv2d(c, 16, kernel_size=4, stride=2)
        self.conv2 = nn.Conv2d(16, 32, kernel_size=4, stride=2, padding=2)
        self.conv3 = nn.Conv2d(32, 256, kernel_size=11, stride=1)
        h1, w1 = (h - 4) // 2 + 1, (w - 4) // 2 + 1
        h2, w2 = (h1 + 4 - 4) // 2 + 1, (w1 + 4 - 4) // 2 + 1
        feats = 256 * (h2 - 11 + 1) * (w2 - 11 + 1)
        self.critic_linear = nn.Linear(feats, 1)
        self.actor_linear = nn.Linear(feats, num_actions)

    def forward(self, x):
        x = F.relu(self.conv1(x / 255.0))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x)).flatten(1)
        return self.actor_linear(x), self.critic_linear(x)


def find_resource_dir(resource_dir=None):
    """Directory holding the reference's checkpoint-*.pkl files, or None."""
    cands = [resource_dir, os.environ.get("COMPETITIVE_RL_RESOURCES")]
    try:
        import competitive_rl   # a stock install of the reference keeps them in <repo>/resources/pong
        cands.append(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(competitive_rl.__file__))), "resources", "pong"))
    except Exception:  # noqa: BLE001
        pass
    for c in cands:
        if c and os.path.isdir(c):
            return c
    return None


class DevicePolicy(object):
    """policy_serving.Policy on the device: keeps the opponent's own stack of the last `frame_stack` single frames
    (FrameStackTensor.update: roll, append; never reset on done, like the reference) and returns greedy actions."""

    def __init__(self, num_envs, checkpoint_path="", use_light_model=True, device="cuda", obs_shape=(1, 42, 42),
                 frame_stack=4, num_actions=3):
        self.num_envs, self.device = num_envs, torch.device(device)
        self.num_channels = obs_shape[0]
        shape = (obs_shape[0] * frame_stack, *obs_shape[1:])
        net = LightActorCritic if use_light_model else ActorCritic
        self.model = net(shape, num_actions).to(self.device)
        if checkpoint_path:
            # the reference's checkpoint-*.pkl files hold {"model": state_dict, ...} of plain tensors: no pickled code is needed
            state = torch.load(checkpoint_path, map_location=self.device, weights_only=True)
            self.model.load_state_dict(state["model"] if "model" in state else state)
        self.model.requires_grad_(False)
        self.stack = torch.zeros((num_envs, *shape), dtype=torch.float32, device=self.device)

    def reset(self):
        self.stack.zero_()

    @torch.no_grad()
    def logits(self, obs):
        obs = torch.as_tensor(obs).to(self.device, dtype=torch.float32).reshape(self.num_envs, self.num_channels, *self.stack.shape[2:])
        self.stack = self.stack.roll(shifts=-self.num_channels, dims=1)
        self.stack[:, -self.num_channels:] = obs
        # plain fp32 convolutions (cuDNN would otherwise take TF32 tensor-core paths on this GPU: the networks are tiny,
        # and a greedy argmax over near-tied logits should not depend on the conv algorithm's precision)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return self.model(self.stack)[0]

    def __call__(self, obs):
        """-> int32 [num_envs] device tensor of greedy actions (Categorical(logits).probs.argmax, policy_serving.py:50-56)."""
        return self.logits(obs).argmax(dim=1).to(torch.int32)


def get_builtin_agent_names(resource_dir=None):
    """Agents that can actually be built here: RANDOM, RULE_BASED and every network agent whose checkpoint exists."""
    d = find_resource_dir(resource_dir)
    names = []
    for n in BUILTIN_AGENT_NAMES:
        if n in _CHECKPOINTS and not (d and os.path.isfile(os.path.join(d, _CHECKPOINTS[n][0]))):
            continue
        names.append(n)
    return names


def get_compute_action_function(agent_name, num_envs=1, device="cuda", resource_dir=None):
    """Callable obs -> int32 [num_envs] device tensor (pong/builtin_policies.py:61-91)."""
    device = torch.device(device)
    if agent_name == "RULE_BASED":
        return lambda _obs: torch.full((num_envs,), CHEAT_CODES, dtype=torch.int32, device=device)
    if agent_name == "RANDOM":
        return lambda _obs: torch.randint(0, 3, (num_envs,), dtype=torch.int32, device=device)
    if agent_name in _CHECKPOINTS:
        d = find_resource_dir(resource_dir)
        fname, light = _CHECKPOINTS[agent_name]
        path = os.path.join(d, fname) if d else None
        if not path or not os.path.isfile(path):
            raise FileNotFoundError("checkpoint %s of built-in agent %s not found; set COMPETITIVE_RL_RESOURCES to the "
                                    "reference's resources/pong directory" % (fname, agent_name))
        return DevicePolicy(num_envs, path, light, device)
    raise ValueError("Unknown agent name: {}".format(agent_name))
