"""Rate-limited actions: optimiser variables are scaled per-step changes, accumulated over the horizon and
clamped to [0,1] with a straight-through gradient (reference actions_mappers/derivative_action_mapper.py:10-35)."""
import torch

from rl_gp_mpc.control_objects.utils.pytorch_utils import Clamp

from .abstract_action_mapper import AbstractActionMapper


class DerivativeActionMapper(AbstractActionMapper):
    def __init__(self, action_low, action_high, len_horizon, config):
        super().__init__(action_low, action_high, len_horizon, config)
        self.action_model_previous_iter = torch.rand(self.dim_action)
        self.bounds = [(0, 1)] * (self.dim_action * len_horizon)
        self.clamp_class = Clamp()

    def transform_action_raw_to_action_model(self, action_raw):
        return self.norm_action(action_raw)

    def transform_action_model_to_action_raw(self, action_model, update_internals: bool = False):
        if update_internals:
            self.action_model_previous_iter = action_model[0]
        return self.denorm_action(action_model, update_internals=update_internals)

    def transform_action_mpc_to_action_model(self, action_mpc):
        max_change = self.config.max_change_action_norm
        steps = torch.atleast_2d(torch.as_tensor(action_mpc).reshape(self.len_horizon, -1)) * (2 * max_change) - max_change
        first = (steps[0] + self.action_model_previous_iter)[None]
        cumulative = torch.cumsum(torch.cat((first, steps[1:]), 0), dim=0)
        return Clamp.apply(cumulative, 0, 1)
