"""Latency of the legacy single-sequence objective (GpMpcController.compute_mean_lcb_trajectory, what scipy's L-BFGS-B
calls once per iteration) and of one get_action control step, on small closed-loop-sized problems."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200"))

from oracle.workloads import make_workload  # noqa: E402
from rl_gp_mpc import GpMpcController  # noqa: E402
from rl_gp_mpc.config_classes.actions_config import ActionsConfig  # noqa: E402
from rl_gp_mpc.config_classes.controller_config import ControllerConfig  # noqa: E402
from rl_gp_mpc.config_classes.model_config import ModelConfig  # noqa: E402
from rl_gp_mpc.config_classes.observation_config import ObservationConfig  # noqa: E402
from rl_gp_mpc.config_classes.reward_config import RewardConfig  # noqa: E402
from rl_gp_mpc.config_classes.total_config import Config  # noqa: E402


def controller_for(cfg):
    E, Na, H = cfg["E"], cfg["Na"], cfg["H"]
    r = cfg["reward"]
    config = Config(
        observation_config=ObservationConfig(obs_var_norm=[cfg["obs_var"]] * E),
        reward_config=RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                                   weight_state_terminal=list(r["weight_state_terminal"]),
                                   target_action_norm=list(r["target_action"]), weight_action=list(r["weight_action"]),
                                   exploration_factor=r["exploration_factor"]),
        actions_config=ActionsConfig(), controller_config=ControllerConfig(len_horizon=H),
        model_config=ModelConfig(gp_init={"noise_covar.noise": list(cfg["noise"]),
                                          "base_kernel.lengthscale": [list(v) for v in cfg["lengthscale"]],
                                          "outputscale": list(cfg["outputscale"])},
                                 min_std_noise=1e-4, max_std_noise=1.0, min_outputscale=1e-6, max_outputscale=10.0,
                                 min_lengthscale=1e-3, max_lengthscale=1e3))
    ctrl = GpMpcController(-np.ones(E), np.ones(E), -np.ones(Na), np.ones(Na), config)
    ctrl.transition_model.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    return ctrl


def main():
    for name, kw in (("C1 N=50 H=15", dict(name="C1")), ("C2 N=200 H=25", dict(name="C2", B=1)),
                     ("C4b N=500 H=30", dict(name="C4b", B=1)),
                     ("C2 N=200 H=25, per-GP hyper-parameters", dict(name="C2", B=1, distinct_lengthscales=True)),
                     ("C4b N=500 H=30, per-GP hyper-parameters", dict(name="C4b", B=1, distinct_lengthscales=True))):
        cfg = make_workload(kw.pop("name"), **kw)
        ctrl = controller_for(cfg)
        a = cfg["actions"][0].reshape(-1)
        mu, var = torch.as_tensor(cfg["mu0"]), torch.as_tensor(cfg["Sigma0"])
        for _ in range(5):
            ctrl.compute_mean_lcb_trajectory(a, mu, var)
        torch.cuda.synchronize()
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            ctrl.compute_mean_lcb_trajectory(a, mu, var)
        dt = (time.perf_counter() - t0) / n
        eng = ctrl.transition_model.engine
        eng.enable_timing(True)
        ctrl.compute_mean_lcb_trajectory(a, mu, var)
        torch.cuda.synchronize()
        print("%-42s objective+gradient call %.3f ms  (kernels: fwd %.3f + reverse %.3f ms)" % (
            name, dt * 1e3, eng.last_rollout_ms(), eng.last_backward_ms()), flush=True)


if __name__ == "__main__":
    main()
