"""Closed-loop control of the ProcessControl tank by GP-MPC on the B200 backend (the reference's
examples/process_control/run_process_control.py:12-37 with the same environment and controller settings).

    python examples/process_control/run_process_control.py --steps 500 --random-init 100
    python examples/process_control/run_process_control.py --steps 60 --random-init 15 --repeat 1 --batched 256

Prints the mean cost of the random phase and of the controlled phase, the time per control step, and which kernel
path / factorisation update was in use."""
import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200"))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

from config_process_control import get_config  # noqa: E402
from rl_gp_mpc.config_classes.visu_config import VisuConfig  # noqa: E402
from rl_gp_mpc.envs.process_control import ProcessControl  # noqa: E402
from rl_gp_mpc.run_env_function import run_env  # noqa: E402


def run_process_control(num_steps=500, random_actions_init=100, num_repeat_actions=5, include_time_model=False,
                        len_horizon=5, verbose=False, training_frequency=15, batched_candidates=0, seed=None,
                        folder_save=None):
    if seed is not None:
        np.random.seed(seed)
    env = ProcessControl(dt=1, s_range=(20, 30), fi_range=(0.15, 0.3), ci_range=(0.15, 0.2), cr_range=(0.8, 1.0),
                         noise_l_prop_range=(5e-3, 1e-2), noise_co_prop_range=(5e-3, 1e-2), sp_l_range=(0.4, 0.6),
                         sp_co_range=(0.4, 0.6), change_params=False, period_change=200)
    control_config = get_config(len_horizon=len_horizon, include_time_model=include_time_model,
                                num_repeat_actions=num_repeat_actions, training_frequency=training_frequency,
                                batched_candidates=batched_candidates)
    visu_config = VisuConfig(render_live_plot_2d=False, render_env=False, save_render_env=False, save_live_plot_2d=False)
    return run_env(env, control_config, visu_config, random_actions_init=random_actions_init, num_steps=num_steps,
                   verbose=verbose, folder_save=folder_save)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--random-init", type=int, default=100)
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--horizon", type=int, default=5)
    ap.add_argument("--time-model", action="store_true")
    ap.add_argument("--training-frequency", type=int, default=15)
    ap.add_argument("--batched", type=int, default=0)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    t0 = time.time()
    costs = run_process_control(a.steps, a.random_init, a.repeat, a.time_model, a.horizon, a.verbose,
                                a.training_frequency, a.batched, a.seed)
    dt = time.time() - t0
    k = min(a.random_init, len(costs))
    print("steps %d (%.2f s, %.1f ms per env step): mean cost random phase %.4f, controlled phase %.4f, last quarter %.4f" % (
        len(costs), dt, 1e3 * dt / max(len(costs), 1), float(np.mean(costs[:k])) if k else float("nan"),
        float(np.mean(costs[k:])) if len(costs) > k else float("nan"), float(np.mean(costs[-max(len(costs) // 4, 1):]))))
