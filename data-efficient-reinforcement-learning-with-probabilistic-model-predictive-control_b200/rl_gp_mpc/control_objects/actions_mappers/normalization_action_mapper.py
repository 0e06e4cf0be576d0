"""Identity mapping between optimiser variables and normalised actions
(reference actions_mappers/normalization_action_mapper.py:10-23)."""
import torch

from .abstract_action_mapper import AbstractActionMapper


class NormalizationActionMapper(AbstractActionMapper):
    def __init__(self, action_low, action_high, len_horizon, config):
        super().__init__(action_low, action_high, len_horizon, config)
        self.bounds = [(0, 1)] * (self.dim_action * len_horizon)

    def transform_action_raw_to_action_model(self, action_raw):
        return self.norm_action(action_raw)

    def transform_action_model_to_action_raw(self, action_model, update_internals: bool = False):
        return self.denorm_action(action_model, update_internals=update_internals)

    def transform_action_mpc_to_action_model(self, action_mpc):
        return torch.atleast_2d(torch.as_tensor(action_mpc).reshape(self.len_horizon, -1))
