"""RewardConfig (reference config_classes/reward_config.py:4-64): quadratic set-point cost description."""
import torch

from .utils.functions_process_config import convert_config_lists_to_tensor


class RewardConfig:
    def __init__(self, target_state_norm: "list[float]" = None, weight_state: "list[float]" = None,
                 weight_state_terminal: "list[float]" = None, target_action_norm: "list[float]" = None,
                 weight_action: "list[float]" = None, exploration_factor: float = 3, use_constraints: bool = False,
                 state_min: "list[float]" = None, state_max: "list[float]" = None, area_multiplier: float = 1,
                 clip_lower_bound_cost_to_0: bool = False):
        def dflt(v, d):
            return d if v is None else v
        self.target_state_norm = dflt(target_state_norm, [1, 0.5, 0.5])
        self.weight_state = dflt(weight_state, [1, 0.1, 0.1])
        self.weight_state_terminal = dflt(weight_state_terminal, [10, 5, 5])
        self.target_action_norm = dflt(target_action_norm, [0.5])
        self.weight_action = dflt(weight_action, [0.05])
        self.exploration_factor = exploration_factor
        self.use_constraints = use_constraints
        self.state_min = dflt(state_min, [-0.1, 0.05, 0.05])
        self.state_max = dflt(state_max, [1.1, 0.95, 0.925])
        self.area_multiplier = area_multiplier
        self.clip_lower_bound_cost_to_0 = clip_lower_bound_cost_to_0
        self.target_state_action_norm = None
        self.weight_matrix_cost = None
        self.weight_matrix_cost_terminal = None
        convert_config_lists_to_tensor(self)
        combine_weight_matrix(self)
        self.target_state_action_norm = torch.cat((self.target_state_norm, self.target_action_norm))


def combine_weight_matrix(reward_config: RewardConfig):
    """Block-diagonal stage weight (state, action) and terminal weight (reference :58-64)."""
    w = torch.cat((reward_config.weight_state, reward_config.weight_action))
    reward_config.weight_matrix_cost = torch.diag(w)
    reward_config.weight_matrix_cost_terminal = torch.diag(reward_config.weight_state_terminal)
    return reward_config
