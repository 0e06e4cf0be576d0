"""ObservationConfig (reference config_classes/observation_config.py:3-12)."""
import torch


class ObservationConfig:
    def __init__(self, obs_var_norm: "list[float]" = None):
        """obs_var_norm: variance of the normalised observation, one entry per state dimension."""
        diag = [1e-6, 1e-6, 1e-6] if obs_var_norm is None else obs_var_norm
        self.obs_var_norm = torch.diag(torch.tensor(diag, dtype=torch.get_default_dtype()))
