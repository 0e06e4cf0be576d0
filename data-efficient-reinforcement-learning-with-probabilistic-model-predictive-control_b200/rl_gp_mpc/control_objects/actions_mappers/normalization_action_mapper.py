"""Optimiser variables ARE the normalised actions (reference actions_mappers/normalization_action_mapper.py:10-23):
the flat (H*Na,) vector is only reshaped to (H, Na); raw <-> model is the affine min-max map of the base class."""
import torch

from .abstract_action_mapper import AbstractActionMapper


class NormalizationActionMapper(AbstractActionMapper):
    def __init__(self, action_low, action_high, len_horizon, config):
        AbstractActionMapper.__init__(self, action_low, action_high, len_horizon, config)
        n_vars = len_horizon * self.dim_action
        self.bounds = [(0, 1) for _ in range(n_vars)]      # box constraints handed to L-BFGS-B

    def transform_action_mpc_to_action_model(self, action_mpc):
        flat = torch.as_tensor(action_mpc)
        return torch.atleast_2d(flat.reshape(self.len_horizon, -1))

    def transform_action_model_to_action_raw(self, action_model, update_internals: bool = False):
        return self.denorm_action(action_model, update_internals=update_internals)

    def transform_action_raw_to_action_model(self, action_raw):
        return self.norm_action(action_raw)
