"""B200-native GP-MPC inner loop behind the reference's `rl_gp_mpc` API (control_objects + config_classes).

`from rl_gp_mpc import GpMpcController, ControlVisualizations` works like the reference's rl_gp_mpc/__init__.py:1-2;
the driver loop is rl_gp_mpc.run_env_function (run_env / run_env_multiple), the custom environment
rl_gp_mpc.envs.process_control.ProcessControl."""
from . import _cabi  # noqa: F401
from .config_classes import total_config as _total_config  # noqa: F401  (float64 default dtype, total_config.py:11)
from .control_objects.controllers.gp_mpc_controller import GpMpcController  # noqa: F401
from .visu_objects.visu_object import ControlVisualizations  # noqa: F401
