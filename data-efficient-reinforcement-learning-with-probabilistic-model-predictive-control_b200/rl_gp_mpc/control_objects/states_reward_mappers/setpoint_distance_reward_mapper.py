"""Expected quadratic set-point cost of a Gaussian state (reference states_reward_mappers/
setpoint_distance_reward_mapper.py:12-68, :124-149).

Inside the MPC hot loop these formulas run fused in the CUDA rollout kernel (csrc/gpmpc_rollout.cu,
stage_cost / terminal_cost); the methods here serve the reference's out-of-loop callers
(GpMpcController.get_action :88, compute_cost_unnormalized :304) with identical semantics, including the
reference's use of the VARIANCE as sigma in the constraint penalty (:60-64) and the unused area_multiplier.
Returns rewards (= -cost) and cost variances."""
import torch

from rl_gp_mpc.config_classes.reward_config import RewardConfig

from ..utils.pytorch_utils import normal_cdf
from .abstract_state_reward_mapper import AbstractStateRewardMapper


def _quadratic_cost(err, var, weight):
    """mean / variance of (x-t)^T W (x-t) for x ~ N(., var); batched over leading dims."""
    ws = weight @ var
    mean = torch.diagonal(var @ weight, dim1=-2, dim2=-1).sum(-1) + torch.einsum("...i,ij,...j->...", err, weight, err)
    variance = 2.0 * torch.diagonal(ws @ ws, dim1=-2, dim2=-1).sum(-1) \
        + 4.0 * torch.einsum("...i,...ij,...j->...", err, ws @ weight, err)
    return mean, variance


class SetpointStateRewardMapper(AbstractStateRewardMapper):
    def __init__(self, config: RewardConfig):
        super().__init__(config)

    def get_reward(self, state_mu, state_var, action):
        cfg = self.config
        n_a = action.shape[-1]
        err = torch.cat((state_mu, action), -1) - cfg.target_state_action_norm
        lead = state_var.shape[:-2]
        e = state_var.shape[-1]
        full = torch.zeros(lead + (e + n_a, e + n_a), dtype=state_var.dtype)
        full[..., :e, :e] = state_var
        cost_mu, cost_var = _quadratic_cost(err, full, cfg.weight_matrix_cost)
        if cfg.use_constraints:
            sigma = torch.diagonal(state_var, dim1=-2, dim2=-1)   # variance passed as sigma, as in the reference
            below = normal_cdf(cfg.state_min, state_mu, sigma)
            above = 1.0 - normal_cdf(cfg.state_max, state_mu, sigma)
            cost_mu = cost_mu + above.sum(-1) + below.sum(-1)
        return -cost_mu, cost_var

    def get_reward_terminal(self, state_mu, state_var):
        err = state_mu - self.config.target_state_norm
        cost_mu, cost_var = _quadratic_cost(err, state_var, self.config.weight_matrix_cost_terminal)
        return -cost_mu, cost_var

    def get_rewards_trajectory(self, states_mu, states_var, actions):
        r, rv = self.get_reward(states_mu[:-1], states_var[:-1], actions)
        r_end, rv_end = self.get_reward_terminal(states_mu[-1], states_var[-1])
        return torch.cat((r, r_end[None]), 0), torch.cat((rv, rv_end[None]), 0)
