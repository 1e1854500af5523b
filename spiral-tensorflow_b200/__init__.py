"""spiral-tensorflow_b200 -- B200-native SPIRAL inner loop (render -> D reward -> A2C loss).

Host-side mirror of the reference's env/model API (envs/, models/, rl_utils, config) over
libspiral_b200.so, a C-ABI library of hand-written sm_100a CUDA kernels.  The directory name
is not a Python identifier; import it through the ``spiral_b200`` alias module at the repo root
(``import spiral_b200 as sp``) or ``importlib.import_module("spiral-tensorflow_b200")``.
"""
from . import _native  # noqa: F401
from . import brush, config  # noqa: F401
from .config import get_args  # noqa: F401

__all__ = ["_native", "brush", "config", "get_args", "create_env"]


def create_env(args, num_envs=None, device=None):
    from .envs import create_env as _create
    return _create(args, num_envs=num_envs, device=device)
