"""Per-control-step record (reference control_objects/controllers/iteration_info_class.py:6-57); same fields."""
import numpy as np
import torch

NUM_DECIMALS_REPR = 3


def _np(v):
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


class IterationInformation:
    def __init__(self, iteration, state, cost, cost_std, mean_predicted_cost, mean_predicted_cost_std,
                 lower_bound_mean_predicted_cost, predicted_idxs, predicted_states, predicted_states_std,
                 predicted_actions, predicted_costs, predicted_costs_std):
        self.iteration = iteration
        self.state = state
        self.cost = cost
        self.cost_std = cost_std
        self.mean_predicted_cost = mean_predicted_cost
        self.mean_predicted_cost_std = mean_predicted_cost_std
        self.lower_bound_mean_predicted_cost = lower_bound_mean_predicted_cost
        self.predicted_idxs = predicted_idxs
        self.predicted_states = predicted_states
        self.predicted_states_std = predicted_states_std
        self.predicted_actions = predicted_actions
        self.predicted_costs = predicted_costs
        self.predicted_costs_std = predicted_costs_std

    def to_arrays(self):
        for key, value in list(vars(self).items()):
            if isinstance(value, torch.Tensor):
                setattr(self, key, _np(value))

    def to_tensors(self):
        for key, value in list(vars(self).items()):
            if isinstance(value, np.ndarray):
                setattr(self, key, torch.as_tensor(value))

    def __str__(self):
        np.set_printoptions(precision=NUM_DECIMALS_REPR, suppress=True)
        return "\n".join("%s: %s" % (k, _np(v) if isinstance(v, (torch.Tensor, np.ndarray)) else v)
                         for k, v in vars(self).items())
