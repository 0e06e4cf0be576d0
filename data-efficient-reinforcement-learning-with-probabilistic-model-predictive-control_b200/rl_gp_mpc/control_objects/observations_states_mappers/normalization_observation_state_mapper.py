"""Min-max normalisation of observations (reference observations_states_mappers/
normalization_observation_state_mapper.py:10-22)."""
import numpy as np
import torch

from .abstract_observation_state_mapper import AbstractObservationStateMapper


class NormalizationObservationStateMapper(AbstractObservationStateMapper):
    def __init__(self, observation_low, observation_high, config):
        super().__init__(observation_low, observation_high, config)

    def get_state(self, obs, obs_var=None, update_internals=False):
        obs_t = torch.as_tensor(np.asarray(obs), dtype=torch.get_default_dtype())
        state = (obs_t - self.obs_low) / (self.obs_high - self.obs_low)
        if obs_var is None:
            return state, self.config.obs_var_norm
        var_t = torch.as_tensor(np.asarray(obs_var), dtype=torch.get_default_dtype())
        return state, var_t / self.var_norm_factor

    def get_obs(self, state, state_var=None):
        obs = torch.as_tensor(state) * (self.obs_high - self.obs_low) + self.obs_low
        if state_var is None:
            return obs, None
        return obs, torch.as_tensor(state_var) * self.var_norm_factor
