"""Device-side helpers a trainer uses next to the vec-env.

FrameStackTensor mirrors competitive_rl/utils/utils.py:145-173: the stacking the reference's own
training code applies to cPongDouble observations (make_envs cannot FrameStack tuples), with its
zero-on-done semantics -- a finished env's whole stack is zeroed before the new frame is appended,
unlike atari_wrappers.FrameStack which refills with the reset frame."""
import numpy as np
import torch


class FrameStackTensor(object):
    def __init__(self, num_envs, obs_shape, frame_stack, device):
        self.num_channels = obs_shape[0]
        self.obs_shape = (obs_shape[0] * frame_stack, *obs_shape[1:])
        self.current_obs = torch.zeros(num_envs, *self.obs_shape, device=device, dtype=torch.float)
        self.mask_shape = [1] * self.current_obs.dim()
        self.mask_shape[0] = -1
        self.device = device

    def reset(self):
        self.current_obs.fill_(0)

    def update(self, obs, mask=None):
        """current_obs is [num_envs, num_stacks, H, W]; rolls along dim 1 to keep the latest frames.
        `mask` is 0 for envs whose episode just ended (their history is cleared), 1 otherwise."""
        if mask is not None:
            mask = torch.as_tensor(mask, dtype=torch.float, device=self.device).reshape(self.mask_shape)
            self.current_obs *= mask
        self.current_obs = self.current_obs.roll(shifts=-self.num_channels, dims=1)
        if not isinstance(obs, torch.Tensor):
            obs = torch.from_numpy(np.asarray(obs).astype(np.float32))
        self.current_obs[:, -self.num_channels:] = obs.to(self.device, dtype=torch.float)
        return self.current_obs

    def get(self):
        return self.current_obs
