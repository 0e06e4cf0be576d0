"""Base class of the action mappers (reference actions_mappers/abstract_action_mapper.py:9-46): the optimiser
works on "mpc" variables in [0,1]; the model on normalised actions in [0,1]; the env on raw actions."""
from typing import Union

import numpy as np
import torch

from rl_gp_mpc.config_classes.actions_config import ActionsConfig


class AbstractActionMapper:
    def __init__(self, action_low: Union[np.ndarray, torch.Tensor], action_high: Union[np.ndarray, torch.Tensor],
                 len_horizon: int, config: ActionsConfig):
        self.config = config
        self.action_low = torch.as_tensor(np.asarray(action_low), dtype=torch.get_default_dtype())
        self.action_high = torch.as_tensor(np.asarray(action_high), dtype=torch.get_default_dtype())
        self.dim_action = len(action_low)
        self.len_horizon = len_horizon
        self.n_iter_ctrl = 0

    def transform_action_raw_to_action_model(self, action_raw):
        raise NotImplementedError

    def transform_action_model_to_action_raw(self, action_model, update_internals: bool = False):
        raise NotImplementedError

    def transform_action_mpc_to_action_model(self, action_mpc):
        raise NotImplementedError

    def transform_action_mpc_to_action_raw(self, action_mpc, update_internals: bool = False):
        model = self.transform_action_mpc_to_action_model(action_mpc)
        return self.transform_action_model_to_action_raw(model, update_internals=update_internals)

    def norm_action(self, action) -> torch.Tensor:
        a = torch.as_tensor(np.asarray(action), dtype=torch.get_default_dtype())
        return (a - self.action_low) / (self.action_high - self.action_low)

    def denorm_action(self, normed_action, update_internals=False) -> torch.Tensor:
        if update_internals:  # the action is about to be applied to the env
            self.n_iter_ctrl += 1
        a = torch.as_tensor(np.asarray(normed_action), dtype=torch.get_default_dtype())
        return a * (self.action_high - self.action_low) + self.action_low
