"""Rate-limited actions (reference actions_mappers/derivative_action_mapper.py:10-35).

The optimiser variables in [0,1] encode per-step CHANGES of the normalised action, each within
+-max_change_action_norm; the model actions are their running sum starting from the action applied at the previous
control step, clamped to [0,1] with a straight-through gradient.  Inside the fused CUDA rollout the same mapping is
applied by the kernels (forward: scale + cumulative sum + clamp; reverse: reverse cumulative sum)."""
import torch

from rl_gp_mpc.control_objects.utils.pytorch_utils import Clamp

from .abstract_action_mapper import AbstractActionMapper


class DerivativeActionMapper(AbstractActionMapper):
    def __init__(self, action_low, action_high, len_horizon, config):
        AbstractActionMapper.__init__(self, action_low, action_high, len_horizon, config)
        self.bounds = [(0, 1) for _ in range(len_horizon * self.dim_action)]
        self.clamp_class = Clamp
        # no action has been applied yet: the reference starts from a random one (derivative_action_mapper.py:13)
        self.action_model_previous_iter = torch.rand(self.dim_action)

    def transform_action_mpc_to_action_model(self, action_mpc):
        half_width = self.config.max_change_action_norm
        changes = torch.atleast_2d(torch.as_tensor(action_mpc).reshape(self.len_horizon, -1))
        changes = changes * (2 * half_width) - half_width
        start = (changes[0] + self.action_model_previous_iter).unsqueeze(0)
        running = torch.cumsum(torch.cat((start, changes[1:]), dim=0), dim=0)
        return Clamp.apply(running, 0, 1)

    def transform_action_model_to_action_raw(self, action_model, update_internals: bool = False):
        if update_internals:      # the first planned action is the one being applied now
            self.action_model_previous_iter = action_model[0]
        return self.denorm_action(action_model, update_internals=update_internals)

    def transform_action_raw_to_action_model(self, action_raw):
        return self.norm_action(action_raw)
