"""Observation -> normalised state mapping base (reference observations_states_mappers/abstract_observation_state_mapper.py)."""
import numpy as np
import torch


class AbstractObservationStateMapper:
    def __init__(self, observation_low, observation_high, config):
        self.config = config
        self.obs_low = torch.as_tensor(np.asarray(observation_low), dtype=torch.get_default_dtype())
        self.obs_high = torch.as_tensor(np.asarray(observation_high), dtype=torch.get_default_dtype())
        # NB: per-dimension span^2 as a VECTOR, broadcast over the last axis of obs_var -- the reference's
        # behaviour (abstract_observation_state_mapper.py:13), kept as is
        self.var_norm_factor = (self.obs_high - self.obs_low) ** 2
        self.dim_observation = len(observation_low)
        self.dim_state = self.dim_observation

    def get_state(self, obs, obs_var=None, update_internals=False):
        raise NotImplementedError

    def get_obs(self, state, state_var=None):
        raise NotImplementedError
