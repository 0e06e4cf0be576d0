"""Import the UNMODIFIED reference package from /root/reference -- TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the GPU box has no /root/reference).  Used
by oracle/make_golden.py to generate tests/golden/*.npz and by the
`needs_reference` CPU tests that pin oracle/gpmpc_oracle.py against the
reference's own code.

The reference cannot be imported as-is here:
  * rl_gp_mpc/__init__.py:2 -> visu_objects/visu_object.py:7 imports gym,
    matplotlib, imageio (absent)     -> MagicMock stand-ins in sys.modules
  * control_objects/models/gp_model.py:5 imports gpytorch (absent)
                                    -> oracle/gpytorch_shim (Gram matrix +
                                       hyper-parameter containers only)
All hot-path arithmetic except the Gram matrix is then the reference's own
torch code.
"""
import os
import sys
import warnings
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("GPMPC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rl_gp_mpc"))


_loaded = None


def load_reference():
    """Returns the imported reference `rl_gp_mpc` package (verbatim code)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ["gym", "gym.core", "gym.spaces", "gym.wrappers", "gym.wrappers.monitoring",
                 "gym.wrappers.monitoring.video_recorder", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.animation", "matplotlib.gridspec", "matplotlib.widgets",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "imageio", "sklearn.neighbors"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpytorch_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    # the product package mirrors the reference's import name (rl_gp_mpc); make
    # sure we get the reference's, under a private alias, without clobbering it
    saved = {k: v for k, v in sys.modules.items() if k == "rl_gp_mpc" or k.startswith("rl_gp_mpc.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import rl_gp_mpc  # noqa: F401  (reference's; sets default dtype float64, total_config.py:11)
            import rl_gp_mpc.control_objects.controllers.gp_mpc_controller as ctrl
            import rl_gp_mpc.control_objects.models.gp_model as gp_model
            import rl_gp_mpc.control_objects.states_reward_mappers.setpoint_distance_reward_mapper as reward
            import rl_gp_mpc.config_classes.total_config as total_config
            from rl_gp_mpc.config_classes import (actions_config, controller_config, memory_config,
                                                  model_config, observation_config, reward_config,
                                                  training_config)
        ref = {
            "pkg": sys.modules["rl_gp_mpc"],
            "ctrl": ctrl, "gp_model": gp_model, "reward": reward, "total_config": total_config,
            "actions_config": actions_config, "controller_config": controller_config,
            "memory_config": memory_config, "model_config": model_config,
            "observation_config": observation_config, "reward_config": reward_config,
            "training_config": training_config,
        }
    finally:
        sys.path.remove(REFERENCE_ROOT)
        ref_mods = {k: v for k, v in sys.modules.items() if k == "rl_gp_mpc" or k.startswith("rl_gp_mpc.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    _loaded = ref
    return ref


def make_reference_controller(cfg, ref=None):
    """Build the reference GpMpcController for a workload dict (oracle/workloads.py)."""
    import numpy as np
    import torch
    ref = ref or load_reference()
    E, Na = cfg["E"], cfg["Na"]
    rc = cfg["reward"]
    config = ref["total_config"].Config(
        observation_config=ref["observation_config"].ObservationConfig(obs_var_norm=[cfg["obs_var"]] * E),
        reward_config=ref["reward_config"].RewardConfig(
            target_state_norm=list(rc["target_state"]), weight_state=list(rc["weight_state"]),
            weight_state_terminal=list(rc["weight_state_terminal"]),
            target_action_norm=list(rc["target_action"]), weight_action=list(rc["weight_action"]),
            exploration_factor=rc["exploration_factor"], use_constraints=rc["use_constraints"],
            state_min=list(rc["state_min"]), state_max=list(rc["state_max"]), area_multiplier=1,
            clip_lower_bound_cost_to_0=rc["clip_lower_bound_cost_to_0"]),
        actions_config=ref["actions_config"].ActionsConfig(
            limit_action_change=cfg["limit_action_change"],
            max_change_action_norm=list(cfg["max_change_action_norm"])),
        model_config=ref["model_config"].ModelConfig(
            gp_init={"noise_covar.noise": list(cfg["noise"]),
                     "base_kernel.lengthscale": [list(r) for r in cfg["lengthscale"]],
                     "outputscale": list(cfg["outputscale"])},
            min_std_noise=1e-4, max_std_noise=1.0, min_outputscale=1e-6, max_outputscale=10.0,
            min_lengthscale=1e-3, max_lengthscale=1e3, min_lengthscale_time=1e-3,
            max_lengthscale_time=1e5, init_lengthscale_time=cfg.get("lengthscale_time", 100.0),
            include_time_model=cfg["include_time_model"]),
        memory_config=ref["memory_config"].MemoryConfig(),
        training_config=ref["training_config"].TrainingConfig(),
        controller_config=ref["controller_config"].ControllerConfig(len_horizon=cfg["H"]),
    )
    c = ref["ctrl"].GpMpcController(observation_low=-np.ones(E), observation_high=np.ones(E),
                                    action_low=-np.ones(Na), action_high=np.ones(Na), config=config)
    c.iter_ctrl = cfg.get("iter_ctrl", 0)
    if cfg["limit_action_change"]:
        c.actions_mapper.action_model_previous_iter = torch.as_tensor(cfg["action_prev"], dtype=torch.float64)
    return c
