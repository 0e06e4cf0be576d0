"""Replay store feeding the GP training block (reference control_objects/memories/gp_memory.py:10-112).

Host-side producer of (x_mem (N,D), y_mem (N,E)); O(1) work per control step, not part of the accelerated
path.  Same public methods and insertion rule (prediction error / uncertainty gate); buffers grow by
`points_batch_memory` (the reference's growth branch, :35-40/:70-71, is unreachable below 1500 points and
mis-calls torch.cat -- here it simply works)."""
import numpy as np
import torch

from rl_gp_mpc.config_classes.memory_config import MemoryConfig
from rl_gp_mpc.control_objects.utils.data_utils import form_model_input


def _grown(t, extra):
    pad = torch.empty((extra,) + tuple(t.shape[1:]), dtype=t.dtype)
    return torch.cat((t, pad), 0)


class Memory:
    def __init__(self, config: MemoryConfig, dim_input, dim_state, include_time_model=False, step_model=1):
        self.config = config
        self.include_time_model = include_time_model
        self.dim_input, self.dim_state, self.step_model = dim_input, dim_state, step_model
        n = config.points_batch_memory
        self.inputs = torch.empty(n, dim_input)
        self.states_next = torch.empty(n, dim_state)
        self.rewards = torch.empty(n)
        self.iter_ctrls = torch.empty(n)
        self.errors = torch.empty(n, dim_state)
        self.stds = torch.empty(n, dim_state)
        self.model_inputs = torch.empty(n, dim_input)
        self.model_targets = torch.empty(n, dim_state)
        self.active_data_mask = np.empty(n, dtype=bool)
        self.len_mem = 0
        self.len_mem_last_processed = 0
        self.len_mem_model = 0

    def add(self, state, action_model, state_next, reward, iter_ctrl=0, **kwargs):
        if self.len_mem + 1 > len(self.inputs):
            n = self.config.points_batch_memory
            for name in ("inputs", "states_next", "rewards", "iter_ctrls", "errors", "stds"):
                setattr(self, name, _grown(getattr(self, name), n))
            self.active_data_mask = np.concatenate((self.active_data_mask, np.empty(n, dtype=bool)))
        k = self.len_mem
        self.inputs[k] = form_model_input(state=state, action_model=action_model, time_idx=iter_ctrl,
                                          include_time_model=self.include_time_model, dim_input=self.dim_input)
        self.states_next[k] = torch.as_tensor(state_next)
        self.rewards[k] = reward
        self.iter_ctrls[k] = iter_ctrl
        keep = True
        if self.config.check_errors_for_storage:
            pred = kwargs.get("predicted_state")
            pred_std = kwargs.get("predicted_state_std")
            if pred is not None:
                err = torch.abs(torch.as_tensor(pred) - torch.as_tensor(state_next))
                keep = bool(torch.any(err > self.config.min_error_prediction_state_for_memory))
                self.errors[k] = err
            else:
                self.errors[k] = np.nan
            if pred_std is not None:
                pred_std = torch.as_tensor(pred_std)
                keep = keep and bool(torch.any(pred_std > self.config.min_prediction_state_std_for_memory))
                self.stds[k] = pred_std
            else:
                self.stds[k] = np.nan
        self.active_data_mask[k] = keep
        self.len_mem += 1

    def prepare_for_model(self):
        """Append the not-yet-processed, gate-passing points to the model arrays (inputs, delta-state)."""
        idx = self.get_indexes_to_process()
        idx = idx[self.active_data_mask[idx]]
        n_add = len(idx)
        while self.len_mem_model + n_add > len(self.model_inputs):
            self.model_inputs = _grown(self.model_inputs, self.config.points_batch_memory)
            self.model_targets = _grown(self.model_targets, self.config.points_batch_memory)
        if n_add:
            x_new, y_new = self.get_memory_by_index(idx)
            self.model_inputs[self.len_mem_model:self.len_mem_model + n_add] = x_new
            self.model_targets[self.len_mem_model:self.len_mem_model + n_add] = y_new
        self.len_mem_model += n_add
        self.len_mem_last_processed = self.len_mem

    def get_memory_total(self):
        return self.get_memory_by_index(self.get_indexes_processed())

    def get_memory_by_index(self, indexes):
        inputs = self.inputs[indexes]
        targets = self.states_next[indexes + self.step_model - 1] - self.inputs[indexes, :self.dim_state]
        return inputs, targets

    def get_indexes_to_process(self):
        return np.arange(self.len_mem_last_processed, self.len_mem, self.step_model)

    def get_indexes_processed(self):
        return np.arange(0, self.len_mem_last_processed, self.step_model)

    def get_mask_model_inputs(self):
        return self.active_data_mask[self.get_indexes_processed()]

    def get(self):
        if self.len_mem_model > 0:
            return self.model_inputs[:self.len_mem_model], self.model_targets[:self.len_mem_model]
        # empty memory: a single all-zero point so that the model can still be evaluated (reference :109-111)
        return torch.zeros((1, self.dim_input)), torch.zeros((1, self.dim_state))
