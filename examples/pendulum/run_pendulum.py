"""Closed-loop control of Pendulum-v0 by GP-MPC on the B200 backend (the reference's examples/pendulum/run_pendulum.py with
the same controller settings: 150 steps, 10 random actions first, horizon 15, actions held 1 steps).  The
environment is gym's when gym is installed, else the stand-in of rl_gp_mpc/envs/classic_control.py.

    python examples/pendulum/run_pendulum.py --steps 150 --random-init 10
    python examples/pendulum/run_pendulum.py --steps 60 --batched 64          # batched on-device action optimiser
"""
import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200"))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

from config_pendulum import get_config  # noqa: E402
from rl_gp_mpc.config_classes.visu_config import VisuConfig  # noqa: E402
from rl_gp_mpc.envs.classic_control import make  # noqa: E402
from rl_gp_mpc.run_env_function import run_env  # noqa: E402


def run_pendulum(num_steps=150, random_actions_init=10, num_repeat_actions=1, len_horizon=15, verbose=False,
        batched_candidates=0, seed=None, folder_save=None):
    if seed is not None:
        np.random.seed(seed)
    env = make("Pendulum-v0", seed=seed)
    control_config = get_config(len_horizon=len_horizon, num_repeat_actions=num_repeat_actions,
                                batched_candidates=batched_candidates)
    visu_config = VisuConfig(render_live_plot_2d=False, render_env=False, save_render_env=False, save_live_plot_2d=False)
    return run_env(env, control_config, visu_config, random_actions_init=random_actions_init, num_steps=num_steps,
                   verbose=verbose, folder_save=folder_save)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--random-init", type=int, default=10)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--horizon", type=int, default=15)
    ap.add_argument("--batched", type=int, default=0)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    t0 = time.time()
    costs = run_pendulum(a.steps, a.random_init, a.repeat, a.horizon, a.verbose, a.batched, a.seed)
    dt = time.time() - t0
    k = min(a.random_init, len(costs))
    print("steps %d (%.2f s, %.1f ms per env step): mean cost random phase %.4f, controlled phase %.4f, last quarter %.4f" % (
        len(costs), dt, 1e3 * dt / max(len(costs), 1), float(np.mean(costs[:k])) if k else float("nan"),
        float(np.mean(costs[k:])) if len(costs) > k else float("nan"), float(np.mean(costs[-max(len(costs) // 4, 1):]))))
