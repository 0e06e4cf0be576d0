"""Driver loop with the reference's entry points (rl_gp_mpc/run_env_function.py:14-71): closed-loop control of a
gym-style environment (reset/step/observation_space/action_space) by GpMpcController, random actions first."""
import time

import numpy as np

from rl_gp_mpc.config_classes.total_config import Config
from rl_gp_mpc.config_classes.visu_config import VisuConfig
from rl_gp_mpc.control_objects.controllers.gp_mpc_controller import GpMpcController
from rl_gp_mpc.visu_objects.visu_object import ControlVisualizations

NUM_DECIMALS_REPR = 3
np.set_printoptions(precision=NUM_DECIMALS_REPR, suppress=True)


def run_env(env, control_config: Config, visu_config: VisuConfig, random_actions_init=10, num_steps=150, verbose=True,
            device=None, folder_save=None):
    """One episode; returns the per-step costs (numpy, length num_steps)."""
    visu_obj = ControlVisualizations(env=env, num_steps=num_steps, control_config=control_config,
                                     visu_config=visu_config, folder_save=folder_save)
    ctrl_obj = GpMpcController(observation_low=env.observation_space.low, observation_high=env.observation_space.high,
                               action_low=env.action_space.low, action_high=env.action_space.high,
                               config=control_config, device=device)
    obs = env.reset()
    if isinstance(obs, tuple):          # gymnasium: (obs, info)
        obs = obs[0]
    for idx_ctrl in range(num_steps):
        action = ctrl_obj.get_action(obs_mu=obs, random=idx_ctrl < random_actions_init)
        iter_info = ctrl_obj.get_iter_info()
        cost, _cost_var = ctrl_obj.compute_cost_unnormalized(obs, action)
        visu_obj.update(obs=obs, reward=-cost, action=action, env=env, iter_info=iter_info)
        stepped = env.step(action)
        obs_new = stepped[0]
        ctrl_obj.add_memory(obs=obs, action=action, obs_new=obs_new, reward=-cost,
                            predicted_state=iter_info.predicted_states[1],
                            predicted_state_std=iter_info.predicted_states_std[1])
        obs = obs_new
        if verbose:
            print(str(iter_info))
    visu_obj.save(ctrl_obj)
    ctrl_obj.check_and_close_processes()
    ctrl_obj.close()                    # a hyper-parameter fit still running is stopped, not abandoned mid-kernel
    if hasattr(env, "close"):
        env.close()
    visu_obj.close()
    return visu_obj.get_costs()


def run_env_multiple(env, env_name, control_config: Config, visu_config: VisuConfig, num_runs, random_actions_init=10,
                     num_steps=150, verbose=True, device=None):
    """num_runs episodes; saves (and returns) mean / std of the cost per control step."""
    runs = []
    for _ in range(num_runs):
        runs.append(run_env(env, control_config, visu_config, random_actions_init, num_steps, verbose=verbose,
                            device=device))
        time.sleep(1)
    runs = np.array(runs)
    mean, std = runs.mean(axis=0), runs.std(axis=0)
    np.savez("multiple_runs_costs_%s.npz" % env_name, costs=runs, mean=mean, std=std)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        steps = np.arange(len(mean))
        fig, ax = plt.subplots(figsize=(10, 5))
        ax.plot(steps, mean)
        ax.fill_between(steps, mean - std, mean + std, alpha=0.4)
        ax.set_title("Costs of multiples %s runs" % env_name)
        ax.set_ylabel("Cost")
        ax.set_xlabel("Env iteration")
        fig.savefig("multiple_runs_costs_%s.png" % env_name)
        plt.close(fig)
    except Exception:                   # noqa: BLE001 - matplotlib is optional
        pass
    return mean, std
