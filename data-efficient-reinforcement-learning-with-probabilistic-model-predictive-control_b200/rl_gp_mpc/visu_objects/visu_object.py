"""ControlVisualizations with the reference's interface (rl_gp_mpc/visu_objects/visu_object.py:20-114): records the
normalised states / actions / rewards and the per-step IterationInformation of a run, and saves them at the end.

The reference's live matplotlib process, the gym video recorder and the 3-D model plots are presentation layers with
no bearing on the accelerated path; here they degrade gracefully: `save()` writes `run_history.npz` always and a
2-D summary figure (states, actions, costs with the predicted bands) when matplotlib is importable; `render_env` calls
env.render(); live plotting / video capture requests are reported once and skipped."""
import copy
import datetime
import os
import warnings

import numpy as np

from rl_gp_mpc.config_classes.total_config import Config
from rl_gp_mpc.config_classes.visu_config import VisuConfig


def get_env_name(env):
    name = getattr(env, "name", None)
    if name is None:
        spec = getattr(env, "spec", None)
        name = getattr(spec, "id", None) or type(env).__name__
    return str(name)


def create_folder_save(env_name, root="folder_save"):
    path = os.path.join(root, env_name, datetime.datetime.now().strftime("%Y_%m_%d_%H_%M_%S"))
    os.makedirs(path, exist_ok=True)
    return path


class ControlVisualizations:
    def __init__(self, env, num_steps: int, control_config: Config, visu_config: VisuConfig, folder_save=None):
        self.control_config, self.visu_config = control_config, visu_config
        self.states, self.actions, self.rewards, self.model_iter_infos = [], [], [], []
        self.env_str = get_env_name(env)
        self.folder_save = folder_save or create_folder_save(self.env_str)
        self.obs_min, self.obs_max = env.observation_space.low, env.observation_space.high
        self.action_min, self.action_max = env.action_space.low, env.action_space.high
        self.num_steps = num_steps
        if visu_config.render_live_plot_2d or visu_config.save_render_env or visu_config.save_live_plot_2d:
            warnings.warn("live 2-D plotting and video capture are not provided by the B200 backend; the run history "
                          "and a summary figure are saved at the end instead", stacklevel=2)
        self.processes_running = True

    def update(self, obs, action, reward, env, iter_info=None):
        self.states.append((np.asarray(obs) - self.obs_min) / (self.obs_max - self.obs_min))
        self.actions.append((np.asarray(action) - self.action_min) / (self.action_max - self.action_min))
        self.rewards.append(float(reward))
        info = copy.deepcopy(iter_info)
        if info is not None:
            info.to_arrays()
        self.model_iter_infos.append(info)
        self.env_render_step(env)

    def env_render_step(self, env):
        if self.visu_config.render_env:
            try:
                env.render()
            except Exception:           # noqa: BLE001 - headless boxes
                pass

    def save_plot_2d(self):
        states, actions, costs = np.array(self.states), np.array(self.actions), self.get_costs()
        np.savez(os.path.join(self.folder_save, "run_history.npz"), states=states, actions=actions, costs=costs,
                 predicted_costs=np.array([np.asarray(i.predicted_costs) for i in self.model_iter_infos if i is not None]))
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:               # noqa: BLE001
            return
        fig, axes = plt.subplots(3, 1, figsize=(10, 9), sharex=True)
        steps = np.arange(len(states))
        for d in range(states.shape[1]):
            axes[0].plot(steps, states[:, d], label="state %d" % d)
        if self.control_config.reward.use_constraints:
            for bound in (self.control_config.reward.state_min, self.control_config.reward.state_max):
                for v in np.asarray(bound).reshape(-1):
                    axes[0].axhline(float(v), color="k", linestyle=":", linewidth=0.8)
        for d in range(actions.shape[1]):
            axes[1].step(steps, actions[:, d], where="post", label="action %d" % d)
        axes[2].plot(steps, costs, label="cost")
        pred = [(i.mean_predicted_cost, i.mean_predicted_cost_std) for i in self.model_iter_infos if i is not None]
        if len(pred) == len(steps):
            m, s = np.array(pred).T
            axes[2].plot(steps, m, label="mean predicted cost")
            axes[2].fill_between(steps, m - s, m + s, alpha=0.3)
        for ax, name in zip(axes, ("normalised states", "normalised actions", "cost")):
            ax.set_ylabel(name)
            ax.legend(loc="upper right")
        axes[2].set_xlabel("control step")
        fig.savefig(os.path.join(self.folder_save, "history.png"))
        plt.close(fig)

    def save_plot_model_3d(self, controller):
        """The reference draws the GP surfaces here; this backend stores what is needed to redraw them."""
        state = controller.transition_model.save_state()
        state.to_arrays()
        np.savez(os.path.join(self.folder_save, "model_state.npz"), inputs=state.inputs, states_change=state.states_change)

    def save(self, controller):
        self.save_plot_2d()
        self.save_plot_model_3d(controller)

    def close(self):
        self.processes_running = False

    def close_running_processes(self):
        self.processes_running = False

    def __exit__(self, *args):
        if self.processes_running:
            self.close()

    def get_costs(self):
        return -np.array(self.rewards)
