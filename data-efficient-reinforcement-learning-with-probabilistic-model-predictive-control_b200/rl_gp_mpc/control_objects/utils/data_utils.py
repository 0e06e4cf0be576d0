"""Model-input assembly (reference control_objects/utils/data_utils.py:4-9)."""
import torch


def form_model_input(state, action_model, time_idx, include_time_model, dim_input):
    parts = [torch.as_tensor(state), torch.as_tensor(action_model)]
    x = torch.empty(dim_input)
    sa = torch.cat(parts)
    x[:sa.shape[0]] = sa
    if include_time_model:
        x[-1] = time_idx
    return x
