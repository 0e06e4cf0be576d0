"""the reference's driver-side surface (ProcessControl environment, ControlVisualizations,
run_env) against this backend.  The environment and the recorder are host code (CPU tests); the closed loop needs the
CUDA engine (gpu test)."""
import os
import sys

import numpy as np
import pytest


def _process_control_config(len_horizon=4, include_time_model=False, num_repeat_actions=1, maxiter=6):
    """examples/process_control/config_process_control.py:11-90 of the reference, shortened optimiser budget."""
    from rl_gp_mpc.config_classes.actions_config import ActionsConfig
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.memory_config import MemoryConfig
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.config_classes.observation_config import ObservationConfig
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.config_classes.total_config import Config
    from rl_gp_mpc.config_classes.training_config import TrainingConfig
    return Config(
        observation_config=ObservationConfig(obs_var_norm=[1e-6, 1e-6]),
        reward_config=RewardConfig(target_state_norm=[0.5, 0.5], weight_state=[1, 1], weight_state_terminal=[1, 1],
                                   target_action_norm=[0, 0], weight_action=[1e-4, 1e-4], exploration_factor=1,
                                   use_constraints=False, state_min=[0.1, 0.3], state_max=[0.9, 0.8], area_multiplier=1,
                                   clip_lower_bound_cost_to_0=False),
        actions_config=ActionsConfig(limit_action_change=False, max_change_action_norm=[0.1, 0.2]),
        model_config=ModelConfig(gp_init={"noise_covar.noise": [1e-5, 1e-5], "base_kernel.lengthscale": [0.25, 0.25],
                                          "outputscale": [5e-2, 5e-2]},
                                 init_lengthscale_time=100, min_std_noise=1e-3, max_std_noise=3e-1, min_outputscale=1e-5,
                                 max_outputscale=0.95, min_lengthscale=5e-2, max_lengthscale=25.0,
                                 include_time_model=include_time_model, min_lengthscale_time=5, max_lengthscale_time=1000),
        memory_config=MemoryConfig(check_errors_for_storage=True, min_error_prediction_state_for_memory=[1e-5, 1e-5],
                                   min_prediction_state_std_for_memory=[3e-3, 3e-3], points_batch_memory=1500),
        training_config=TrainingConfig(lr_train=7e-3, iter_train=15, training_frequency=10 ** 6, clip_grad_value=1e-3),
        controller_config=ControllerConfig(
            len_horizon=len_horizon,
            actions_optimizer_params={"disp": None, "maxcor": 15, "ftol": 1e-99, "gtol": 1e-99, "eps": 1e-2,
                                      "maxfun": maxiter, "maxiter": maxiter, "iprint": -1, "maxls": 15,
                                      "finite_diff_rel_step": None},
            init_from_previous_actions=True, restarts_optim=1, optimize=True, num_repeat_actions=num_repeat_actions))


def test_process_control_environment_follows_the_tank_balance():
    from rl_gp_mpc.envs.process_control import ProcessControl
    np.random.seed(0)
    env = ProcessControl(noise_l_prop_range=(1e-9, 1e-8), noise_co_prop_range=(1e-9, 1e-8), change_params=False)
    obs = env.reset()
    assert obs.shape == (2,) and env.observation_space.low.shape == (2,) and env.action_space.high.shape == (2,)
    assert 0.3 * 10 - 1e-3 <= obs[0] <= 0.7 * 10 + 1e-3 and 0.3 - 1e-3 <= obs[1] <= 0.7 + 1e-3      # reset window
    v0, r0 = env.v, env.r
    action = np.array([0.3, 0.8])
    obs2, reward, done, info = env.step(action)
    assert env.v == pytest.approx(v0 + (env.fi + 0.8 - 0.3) * env.dt)                                   # volume balance
    assert env.r == pytest.approx(r0 + (env.fi * env.ci + 0.8 * env.cr - 0.3 * r0 / (v0 + 1e-3)) * env.dt)
    assert obs2[0] == pytest.approx(env.v / env.s, abs=1e-5) and obs2[1] == pytest.approx(env.r / env.v, abs=1e-5)
    assert reward == pytest.approx(-((env.v / env.s - env.sp_l) ** 2 + (env.r / env.v - env.sp_co) ** 2), abs=1e-6)
    assert done == 0 and info == {}
    for _ in range(200):                                                                                # stays inside its box
        o, _, _, _ = env.step(env.action_space.sample())
        assert np.all(o >= env.observation_space.low) and np.all(o <= env.observation_space.high)
    env2 = ProcessControl(change_params=True, period_change=3)
    env2.reset()
    s_before = env2.s
    for _ in range(3):
        env2.step(np.array([0.5, 0.5]))
    assert env2.s != s_before                                                                           # parameters redrawn


def test_control_visualizations_records_and_saves(tmp_path):
    from rl_gp_mpc import ControlVisualizations
    from rl_gp_mpc.config_classes.visu_config import VisuConfig
    from rl_gp_mpc.control_objects.controllers.iteration_info_class import IterationInformation
    from rl_gp_mpc.envs.process_control import ProcessControl
    import torch
    np.random.seed(1)
    env = ProcessControl(change_params=False)
    visu = ControlVisualizations(env, num_steps=3, control_config=_process_control_config(),
                                 visu_config=VisuConfig(save_render_env=False, render_live_plot_2d=False, render_env=True),
                                 folder_save=str(tmp_path))
    for k in range(3):
        info = IterationInformation(iteration=k, state=torch.zeros(2), cost=0.1, cost_std=0.01, mean_predicted_cost=0.2,
                                    mean_predicted_cost_std=0.02, lower_bound_mean_predicted_cost=0.1,
                                    predicted_idxs=np.arange(5), predicted_states=torch.zeros(5, 2),
                                    predicted_states_std=torch.zeros(5, 2), predicted_actions=torch.zeros(4, 2),
                                    predicted_costs=torch.zeros(5), predicted_costs_std=torch.zeros(5))
        visu.update(obs=np.array([5.0, 0.5]), action=np.array([0.25, 0.75]), reward=-0.1 * (k + 1), env=env, iter_info=info)
    np.testing.assert_allclose(visu.states[0], [0.5, 0.5])
    np.testing.assert_allclose(visu.actions[0], [0.25, 0.75])
    np.testing.assert_allclose(visu.get_costs(), [0.1, 0.2, 0.3])
    assert isinstance(visu.model_iter_infos[0].predicted_states, np.ndarray)      # deep-copied and converted
    visu.save_plot_2d()
    hist = np.load(os.path.join(str(tmp_path), "run_history.npz"))
    assert hist["states"].shape == (3, 2) and hist["predicted_costs"].shape == (3, 5)
    visu.close()
    assert not visu.processes_running


@pytest.mark.gpu
@pytest.mark.parametrize("include_time_model", [False, True])
def test_run_env_closes_the_loop_on_process_control(tmp_path, include_time_model):
    """run_env: random actions, then MPC steps whose every objective evaluation is a CUDA rollout; the training set grows
    through the O(N^2) append path."""
    import torch
    from rl_gp_mpc.config_classes.visu_config import VisuConfig
    from rl_gp_mpc.envs.process_control import ProcessControl
    from rl_gp_mpc.run_env_function import run_env
    from rl_gp_mpc.control_objects.models import gp_model
    np.random.seed(3)
    torch.manual_seed(3)
    modes = []
    orig = gp_model.GpStateTransitionModel.prepare_inference

    def spy(self, inputs, state_changes):
        orig(self, inputs, state_changes)
        modes.append((self.last_prepare_mode, len(inputs)))
    gp_model.GpStateTransitionModel.prepare_inference = spy
    try:
        env = ProcessControl(change_params=False)
        costs = run_env(env, _process_control_config(include_time_model=include_time_model),
                        VisuConfig(save_render_env=False, render_live_plot_2d=False, render_env=False),
                        random_actions_init=6, num_steps=12, verbose=False, folder_save=str(tmp_path))
    finally:
        gp_model.GpStateTransitionModel.prepare_inference = orig
    assert costs.shape == (12,) and np.all(np.isfinite(costs)) and np.all(costs >= 0)
    assert os.path.isfile(os.path.join(str(tmp_path), "run_history.npz"))
    assert any(m == "append" for m, _ in modes), modes          # the memory grew one point at a time
    assert max(n for _, n in modes) >= 6


def test_classic_control_stand_ins_follow_the_published_dynamics():
    """Pendulum-v0 / MountainCarContinuous-v0 stand-ins (used by examples/pendulum, examples/mountain_car without gym)."""
    from rl_gp_mpc.envs.classic_control import MountainCarContinuous, Pendulum
    p = Pendulum(seed=0)
    obs = p.reset()
    assert obs.shape == (3,) and abs(obs[0] ** 2 + obs[1] ** 2 - 1.0) < 1e-12
    assert np.all(p.observation_space.high == np.array([1, 1, 8], np.float32)) and p.action_space.high[0] == 2
    p.state = np.array([np.pi, 0.0])                       # hanging down, at rest, no torque: stays there
    obs, rew, done, _ = p.step([0.0])
    np.testing.assert_allclose(p.state, [np.pi, 0.0], atol=1e-12)
    assert not done and abs(rew + np.pi ** 2) < 1e-12
    p.state = np.array([0.5, 1.0])                         # one Euler step by hand, torque clipped at 2
    obs, rew, _, _ = p.step([5.0])
    thdot = 1.0 + (-15.0 * np.sin(0.5 + np.pi) + 3.0 * 2.0) * 0.05
    np.testing.assert_allclose(p.state, [0.5 + thdot * 0.05, thdot], atol=1e-12)
    assert abs(rew + (0.25 + 0.1 + 0.001 * 4.0)) < 1e-12
    np.testing.assert_allclose(obs, [np.cos(p.state[0]), np.sin(p.state[0]), thdot], atol=1e-12)
    p.state = np.array([0.0, 7.99])
    p.step([2.0])
    assert p.state[1] <= 8.0                               # speed limit
    m = MountainCarContinuous(seed=0)
    obs = m.reset()
    assert -0.6 <= obs[0] <= -0.4 and obs[1] == 0.0
    m.state = np.array([-0.5, 0.0])
    obs, rew, done, _ = m.step([1.0])
    v = 0.0015 - 0.0025 * np.cos(-1.5)
    np.testing.assert_allclose(obs, [-0.5 + v, v], atol=1e-15)
    assert not done and abs(rew + 0.1) < 1e-15
    m.state = np.array([-1.2, -0.05])                      # inelastic left wall
    obs, _, _, _ = m.step([-1.0])
    assert obs[0] == -1.2 and obs[1] == 0.0
    m.state = np.array([0.44, 0.07])
    obs, rew, done, _ = m.step([1.0])
    assert done and abs(rew - 99.9) < 1e-12
    m.state = np.array([-0.5, 0.0])                        # full throttle alone never climbs out (the point of the task)
    top = max(m.step([1.0])[0][0] for _ in range(300))
    assert top < 0.45


def test_example_configs_carry_the_reference_values():
    """examples/pendulum and examples/mountain_car: dimensions and hyper-parameters of BASELINE.json configs C1-C3."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel, dims, horizon, maxfun in (("examples/pendulum/config_pendulum.py", 3, 15, 4),
                                       ("examples/mountain_car/config_mountaincar.py", 2, 10, 8)):
        spec = importlib.util.spec_from_file_location("cfg_" + str(dims), os.path.join(root, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        cfg = mod.get_config()
        assert cfg.controller.len_horizon == horizon and cfg.controller.actions_optimizer_params["maxfun"] == maxfun
        assert len(cfg.model.gp_init["noise_covar.noise"]) == dims
        assert float(np.asarray(cfg.model.gp_init["outputscale"]).reshape(-1)[0]) == 5e-2
        assert float(np.asarray(cfg.model.gp_init["base_kernel.lengthscale"]).reshape(-1)[0]) == 0.5
        assert cfg.reward.target_state_norm.shape[-1] == dims


@pytest.mark.gpu
@pytest.mark.parametrize("example", ["pendulum", "mountain_car"])
def test_examples_close_the_loop_on_the_stand_in_environments(tmp_path, example):
    import importlib.util
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    name = {"pendulum": "run_pendulum", "mountain_car": "run_mountaincar"}[example]
    sys.path.insert(0, os.path.join(root, "examples", example))
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "examples", example, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        np.random.seed(5)
        torch.manual_seed(5)
        costs = getattr(mod, name)(num_steps=14, random_actions_init=8, num_repeat_actions=1, len_horizon=6, seed=5,
                                   folder_save=str(tmp_path))
    finally:
        sys.path.pop(0)
    assert costs.shape == (14,) and np.all(np.isfinite(costs)) and np.all(costs >= 0)
