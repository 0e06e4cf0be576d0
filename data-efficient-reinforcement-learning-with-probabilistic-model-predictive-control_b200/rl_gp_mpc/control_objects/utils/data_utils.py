"""Assembly of one GP input row [state, action, (time)] (reference control_objects/utils/data_utils.py:4-9)."""
import torch


def form_model_input(state, action_model, time_idx, include_time_model, dim_input):
    row = torch.empty(dim_input)
    n_state = len(state)
    n_sa = n_state + len(action_model)
    row[:n_state] = torch.as_tensor(state)
    row[n_state:n_sa] = torch.as_tensor(action_model)
    if include_time_model:
        row[dim_input - 1] = time_idx      # un-normalised control-step index
    return row
