"""Config post-processing helpers; same names/behaviour as the reference's
rl_gp_mpc/config_classes/utils/functions_process_config.py:4-36 (lists -> float64 tensors, per-dimension broadcast)."""
import torch


def _as_tensor(value):
    return value if isinstance(value, torch.Tensor) else torch.tensor(value, dtype=torch.get_default_dtype())


def convert_config_lists_to_tensor(self):
    for key in list(vars(self)):
        if isinstance(getattr(self, key), list):
            setattr(self, key, _as_tensor(getattr(self, key)))
    return self


def convert_dict_lists_to_dict_tensor(dict_list):
    for key in list(dict_list):
        if isinstance(dict_list[key], list):
            dict_list[key] = _as_tensor(dict_list[key])
    return dict_list


def extend_dim(orig_tensor, dim):
    """Broadcast a scalar / per-model vector to `dim` (reference :29-36)."""
    t = _as_tensor(orig_tensor)
    target = torch.ones(dim)
    if t.ndim < target.ndim:
        t = t.unsqueeze(-1)
    return t * target


def extend_dim_lengthscale_time(lengthscale, lengthscale_time, num_models, num_inputs):
    """(num_models, num_inputs) lengthscales whose last column is the time lengthscale (reference :18-27)."""
    out = torch.empty((num_models, num_inputs))
    out[:, -1] = lengthscale_time
    per_model_vector = isinstance(lengthscale, torch.Tensor) and lengthscale.dim() == 1
    if per_model_vector:
        out[:, :-1] = lengthscale.reshape(-1, 1).expand(num_models, num_inputs - 1)
    else:
        out[:, :-1] = lengthscale
    return out
