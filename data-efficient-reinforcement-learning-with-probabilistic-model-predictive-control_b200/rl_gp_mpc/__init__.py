"""B200-native GP-MPC inner loop behind the reference's `rl_gp_mpc` API (control_objects + config_classes).

`from rl_gp_mpc import GpMpcController` works like the reference's rl_gp_mpc/__init__.py:1; the visualisation
objects (rl_gp_mpc/__init__.py:2) are outside the accelerated path and not provided."""
from . import _cabi  # noqa: F401
from .config_classes import total_config as _total_config  # noqa: F401  (float64 default dtype, total_config.py:11)
from .control_objects.controllers.gp_mpc_controller import GpMpcController  # noqa: F401
