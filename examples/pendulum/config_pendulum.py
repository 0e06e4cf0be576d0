"""Controller configuration of the Pendulum example: the values of the reference's
examples/pendulum/config_pendulum.py:11-96 (observation (cos th, sin th, th_dot), one torque; the BASELINE.json
configs C1 / C2 use these dimensions and hyper-parameters)."""
from rl_gp_mpc.config_classes.actions_config import ActionsConfig
from rl_gp_mpc.config_classes.controller_config import ControllerConfig
from rl_gp_mpc.config_classes.memory_config import MemoryConfig
from rl_gp_mpc.config_classes.model_config import ModelConfig
from rl_gp_mpc.config_classes.observation_config import ObservationConfig
from rl_gp_mpc.config_classes.reward_config import RewardConfig
from rl_gp_mpc.config_classes.total_config import Config
from rl_gp_mpc.config_classes.training_config import TrainingConfig

E = 3                                   # state dimensions
LBFGSB_OPTIONS = {"disp": None, "maxcor": 4, "ftol": 1e-15, "gtol": 1e-15, "eps": 1e-2, "maxfun": 4, "maxiter": 4,
                  "iprint": -1, "maxls": 4, "finite_diff_rel_step": None}


def get_config(len_horizon=15, include_time_model=False, num_repeat_actions=1, batched_candidates=0):
    return Config(
        observation_config=ObservationConfig(obs_var_norm=[1e-6] * E),
        reward_config=RewardConfig(
            target_state_norm=[1, 0.5, 0.5], weight_state=[1, 0.1, 0.1], weight_state_terminal=[5, 2, 2],
            target_action_norm=[0.5], weight_action=[1e-3], exploration_factor=1, use_constraints=False,
            state_min=[-3] * E, state_max=[3] * E, area_multiplier=1, clip_lower_bound_cost_to_0=False),
        actions_config=ActionsConfig(limit_action_change=False, max_change_action_norm=[0.3]),
        model_config=ModelConfig(
            gp_init={"noise_covar.noise": [1e-5] * E, "base_kernel.lengthscale": [0.5] * E, "outputscale": [5e-2] * E},
            init_lengthscale_time=100, min_std_noise=1e-3, max_std_noise=1e-2, min_outputscale=1e-2,
            max_outputscale=0.95, min_lengthscale=4e-3, max_lengthscale=10.0, include_time_model=include_time_model,
            min_lengthscale_time=10, max_lengthscale_time=10000),
        memory_config=MemoryConfig(check_errors_for_storage=True, min_error_prediction_state_for_memory=[3e-4] * E,
                                   min_prediction_state_std_for_memory=[3e-3] * E, points_batch_memory=1500),
        training_config=TrainingConfig(lr_train=7e-3, iter_train=15, training_frequency=25, clip_grad_value=1e-3,
                                       print_train=False, step_print_train=5),
        controller_config=ControllerConfig(len_horizon=len_horizon, actions_optimizer_params=dict(LBFGSB_OPTIONS),
                                           num_repeat_actions=num_repeat_actions,
                                           batched_candidates=batched_candidates))
