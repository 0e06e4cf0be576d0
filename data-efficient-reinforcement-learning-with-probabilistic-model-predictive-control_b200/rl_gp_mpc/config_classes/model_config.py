"""ModelConfig (reference config_classes/model_config.py:4-67): GP hyper-parameter init values and bounds."""
from .utils.functions_process_config import convert_dict_lists_to_dict_tensor, extend_dim, extend_dim_lengthscale_time


class ModelConfig:
    def __init__(self, gp_init: dict = None, init_lengthscale_time: float = 100, min_std_noise: float = 1e-3,
                 max_std_noise: float = 3e-1, min_outputscale: float = 1e-5, max_outputscale: float = 0.95,
                 min_lengthscale: float = 4e-3, max_lengthscale: float = 25.0, min_lengthscale_time: float = 10,
                 max_lengthscale_time: float = 10000, include_time_model: bool = False):
        """gp_init keys: 'noise_covar.noise' (variance), 'base_kernel.lengthscale', 'outputscale' -- one entry per
        state dimension (one GP per state).  include_time_model adds the control-step index as an extra input."""
        if gp_init is None:
            gp_init = {"noise_covar.noise": [1e-4, 1e-4, 1e-4],
                       "base_kernel.lengthscale": [[0.75] * 4, [0.75] * 4, [0.75] * 4],
                       "outputscale": [5e-2, 5e-2, 5e-2]}
        self.include_time_model = include_time_model
        self.min_std_noise, self.max_std_noise = min_std_noise, max_std_noise
        self.min_outputscale, self.max_outputscale = min_outputscale, max_outputscale
        self.min_lengthscale, self.max_lengthscale = min_lengthscale, max_lengthscale
        self.min_lengthscale_time, self.max_lengthscale_time = min_lengthscale_time, max_lengthscale_time
        self.init_lengthscale_time = init_lengthscale_time
        self.gp_init = convert_dict_lists_to_dict_tensor(gp_init)

    def extend_dimensions_params(self, dim_state, dim_input):
        per_model = (dim_state,)
        for name in ("min_std_noise", "max_std_noise", "min_outputscale", "max_outputscale"):
            setattr(self, name, extend_dim(getattr(self, name), dim=per_model))
        for key in ("noise_covar.noise", "outputscale"):
            self.gp_init[key] = extend_dim(self.gp_init[key], dim=per_model)
        ls_key = "base_kernel.lengthscale"
        if self.include_time_model:
            self.min_lengthscale = extend_dim_lengthscale_time(self.min_lengthscale, self.min_lengthscale_time,
                                                               dim_state, dim_input)
            self.max_lengthscale = extend_dim_lengthscale_time(self.max_lengthscale, self.max_lengthscale_time,
                                                               dim_state, dim_input)
            self.gp_init[ls_key] = extend_dim_lengthscale_time(lengthscale=self.gp_init[ls_key],
                                                               lengthscale_time=self.init_lengthscale_time,
                                                               num_models=dim_state, num_inputs=dim_input)
        else:
            full = (dim_state, dim_input)
            self.min_lengthscale = extend_dim(self.min_lengthscale, dim=full)
            self.max_lengthscale = extend_dim(self.max_lengthscale, dim=full)
            self.gp_init[ls_key] = extend_dim(self.gp_init[ls_key], dim=full)
