"""Stand-in for the `gym` API surface the reference's Pong path uses -- TEST
INFRASTRUCTURE ONLY (oracle side).  gym is an un-vendored, unpinned dependency
of the reference (setup.py:6-15) and is not installed in this image.

Restated third-party behaviour (not pinned by /root/reference):
  * old-gym underscore methods: Env.step/reset/seed/render call `_step/_reset/
    _seed/_render` (pong/base_pong_env.py:38-64 defines only those).
  * `spaces.Box` without a dtype: gym's default is float32.  The north star
    specifies the uint8 observation path, so this shim defaults to uint8 when
    `high == 255`; set GYM_SHIM_BOX_FLOAT32=1 to get gym's float32.
"""
from . import envs, error, logger, spaces, utils, wrappers  # noqa: F401
from .core import Env, ObservationWrapper, RewardWrapper, Wrapper  # noqa: F401
from .envs.registration import make, register, spec  # noqa: F401

__version__ = "0.0-shim"
