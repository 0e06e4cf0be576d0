"""CPU oracle for the GP-MPC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 torch/CPU restatement of the reference's algorithm for the path named
by BASELINE.json `north_star`.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it; the product path
(the CUDA extension) never does.

Pinned by: tests/golden/*.npz, produced by oracle/make_golden.py from the
reference's own code imported verbatim from /root/reference (oracle/
ref_loader.py; gpytorch replaced by oracle/gpytorch_shim, whose only arithmetic
is the RBF Gram matrix).  tests/test_oracle_vs_golden.py checks every function
here against those vectors.  The reference itself has no tests / golden vectors
for this path ("parity unpinned" by the reference); and the
Gram matrix follows gpytorch's published formula, not an executed gpytorch.

Each function cites the reference lines it restates (paths relative to
/root/reference/rl_gp_mpc/).  All arithmetic is float64 (config_classes/
total_config.py:11).
"""
import math

import numpy as np
import torch

F64 = torch.float64


def _t(a):
    return torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=F64)


class OracleModel:
    """Factorised GP state (control_objects/models/gp_model.py:182-191)."""

    def __init__(self, x, y, lengthscale, outputscale, noise):
        self.x_mem = _t(x)                      # (N, D)
        self.y_mem = _t(y)                      # (N, E)
        self.lengthscales = _t(lengthscale)     # (E, D)
        self.variances = _t(outputscale)        # (E,)
        self.noise = _t(noise)                  # (E,)
        self.iK, self.beta = calculate_factorizations(self.x_mem, self.y_mem, self.lengthscales,
                                                      self.variances, self.noise)
        self.iL = torch.diag_embed(1.0 / self.lengthscales)  # gp_model.py:191
        self.dim_state = self.y_mem.shape[1]
        self.dim_input = self.x_mem.shape[1]


def gram_matrix(x, lengthscale, outputscale):
    """ScaleKernel(RBFKernel(ard)) Gram matrices, one per GP (gp_model.py:391-392, :425).

    K_a[i,j] = s2_a exp(-1/2 sum_d ((x_i-x_j)_d / l_{a,d})^2); diagonal exactly s2_a.
    """
    xs = x[None, :, :] / lengthscale[:, None, :]               # (E, N, D)
    d2 = ((xs[:, :, None, :] - xs[:, None, :, :]) ** 2).sum(-1)
    return outputscale[:, None, None] * torch.exp(-0.5 * d2)


def calculate_factorizations(x, y, lengthscale, outputscale, noise):
    """iK = (K + s_n^2 I)^-1 and beta = iK y via Cholesky (gp_model.py:400-431)."""
    K = gram_matrix(x, lengthscale, outputscale)
    n = K.shape[1]
    eye = torch.eye(n, dtype=F64).expand(K.shape[0], n, n)
    chol = torch.linalg.cholesky(K + noise[:, None, None] * eye)          # :426-427
    iK = torch.cholesky_solve(eye, chol)                                  # :428
    beta = torch.cholesky_solve(y.t()[:, :, None], chol)[:, :, 0]         # :429-430
    return iK, beta


def predict_next_state_change(model, input_mu, input_var):
    """Moment-matched GP prediction at a Gaussian input (gp_model.py:112-180).

    Returns (M^T (1,E), S (E,E), V^T (D,E)) exactly like the reference.
    """
    E, D = model.dim_state, model.dim_input
    ls, s2 = model.lengthscales, model.variances
    nu = model.x_mem - input_mu                                   # (N, D)     :138
    iN = nu[None, :, :] / ls[:, None, :]                          # (E, N, D)  :140
    eyeD = torch.eye(D, dtype=F64)
    Bm = input_var[None] / (ls[:, :, None] * ls[:, None, :]) + eyeD           # :141
    tt = torch.linalg.solve(Bm, iN.transpose(-1, -2)).transpose(-1, -2)        # :145-146
    lb = torch.exp(-0.5 * (iN * tt).sum(-1)) * model.beta                      # :148
    til = tt / ls[:, None, :]                                                  # :149
    c = s2 / torch.sqrt(torch.det(Bm))                                         # :150
    M = lb.sum(-1) * c                                                         # :152
    V = torch.einsum("and,an->ad", til, lb) * c[:, None]                       # :153
    # predictive covariance                                                      :155-178
    w = 1.0 / ls ** 2                                                          # (E, D)
    R = input_var[None, None] * (w[:, None, None, :] + w[None, :, None, :]) + eyeD   # :156-159
    X = nu[None, None, :, :] * w[:, None, None, :]                             # (E,1,N,D)  :161
    X2 = -nu[None, None, :, :] * w[None, :, None, :]                           # (1,E,N,D)  :162
    Q = 0.5 * torch.linalg.solve(R, input_var.expand(E, E, D, D))              # :163
    X = X.expand(E, E, -1, -1)
    X2 = X2.expand(E, E, -1, -1)
    XQ = X @ Q
    Xs = (XQ * X).sum(-1)                                                      # :164
    X2s = ((X2 @ Q) * X2).sum(-1)                                              # :165
    maha = -2.0 * XQ @ X2.transpose(-1, -2) + Xs[..., :, None] + X2s[..., None, :]   # :166
    k = torch.log(s2)[:, None] - 0.5 * (iN ** 2).sum(-1)                       # :168
    L = torch.exp(k[:, None, :, None] + k[None, :, None, :] + maha)            # :169
    S = torch.einsum("ai,abij,bj->ab", model.beta, L, model.beta)              # :170-171
    diagL = torch.stack([L[a, a] for a in range(E)])                           # :173-174
    S = S - torch.diag_embed((model.iK * diagL).sum((1, 2)))                   # :175
    S = S / torch.sqrt(torch.det(R))                                           # :176
    S = S + torch.diag_embed(s2)                                               # :177
    S = S - M[:, None] * M[None, :]                                            # :178
    return M[None, :], S, V.t()                                                # :180


def predict_trajectory(model, actions, obs_mu, obs_var, len_horizon, current_time_idx=0,
                       include_time_model=False):
    """H-step moment-matching recurrence (gp_model.py:60-110)."""
    E, D = model.dim_state, model.dim_input
    Na = actions.shape[1]
    mus, vars_ = [obs_mu], [obs_var]
    for t in range(1, len_horizon + 1):
        input_var = torch.zeros((D, D), dtype=F64)
        input_var = torch.cat([torch.cat([vars_[-1], torch.zeros((E, D - E), dtype=F64)], 1),
                               torch.zeros((D - E, D), dtype=F64)], 0)         # :96-97
        parts = [mus[-1], actions[t - 1]]
        if include_time_model:
            parts.append(torch.tensor([float(current_time_idx + t - 1)], dtype=F64))   # :101-102
        input_mean = torch.cat(parts)
        dM, S, v = predict_next_state_change(model, input_mean, input_var)     # :103
        mus.append(mus[-1] + dM[0])                                            # :105
        sv = input_var[:E] @ v                                                 # (E, E)
        vars_.append(S + vars_[-1] + sv + sv.t())                              # :106-108
    return torch.stack(mus), torch.stack(vars_)


def normal_cdf(x, mu, sigma):
    """control_objects/utils/pytorch_utils.py:16-17."""
    return 0.5 * (1.0 + torch.erf((x - mu) / (sigma * math.sqrt(2.0))))


class Clamp(torch.autograd.Function):
    """Straight-through clamp (control_objects/utils/pytorch_utils.py:4-13)."""

    @staticmethod
    def forward(ctx, inp, lo, hi):
        return inp.clamp(min=lo, max=hi)

    @staticmethod
    def backward(ctx, g):
        return g.clone(), None, None


class OracleReward:
    """SetpointStateRewardMapper (states_reward_mappers/setpoint_distance_reward_mapper.py)."""

    def __init__(self, reward_cfg):
        r = reward_cfg
        self.target_state = _t(r["target_state"])
        self.target_sa = torch.cat([self.target_state, _t(r["target_action"])])          # reward_config.py:55
        self.W = torch.diag(torch.cat([_t(r["weight_state"]), _t(r["weight_action"])]))  # :58-62
        self.WT = torch.diag(_t(r["weight_state_terminal"]))                              # :63
        self.kappa = float(r["exploration_factor"])
        self.use_constraints = bool(r["use_constraints"])
        self.state_min = _t(r["state_min"])
        self.state_max = _t(r["state_max"])
        self.clip = bool(r["clip_lower_bound_cost_to_0"])

    def get_reward(self, state_mu, state_var, action):
        """Stage reward for (H,E),(H,E,E),(H,Na) or a single (E,),(E,E),(Na,) (:12-68)."""
        single = state_mu.ndim == 1
        if single:
            state_mu, state_var, action = state_mu[None], state_var[None], action[None]
        T, E = state_mu.shape
        Na = action.shape[1]
        err = torch.cat([state_mu, action], -1) - self.target_sa                       # :36
        sav = torch.zeros((T, E + Na, E + Na), dtype=F64)
        sav = torch.cat([torch.cat([state_var, torch.zeros((T, E, Na), dtype=F64)], 2),
                         torch.zeros((T, Na, E + Na), dtype=F64)], 1)                  # :37-44
        cost_mu = torch.einsum("tii->t", sav @ self.W) + torch.einsum("ti,ij,tj->t", err, self.W, err)  # :47-51
        TS = self.W @ sav                                                              # :52
        cost_var = 2.0 * torch.einsum("tii->t", TS @ TS) \
            + 4.0 * torch.einsum("ti,tij,tj->t", err, TS @ self.W, err)                # :53-56
        if self.use_constraints:                                                        # :58-66
            sig = torch.diagonal(state_var, dim1=-2, dim2=-1)   # NB: variance used as sigma (:60-64)
            pmin = normal_cdf(self.state_min, state_mu, sig)
            pmax = 1.0 - normal_cdf(self.state_max, state_mu, sig)
            cost_mu = cost_mu + pmax.sum(-1) + pmin.sum(-1)
        if single:
            return -cost_mu[0], cost_var[0]
        return -cost_mu, cost_var

    def get_reward_terminal(self, state_mu, state_var):
        """:124-142."""
        err = state_mu - self.target_state
        cost_mu = torch.trace(state_var @ self.WT) + err @ self.WT @ err
        TS = self.WT @ state_var
        cost_var = torch.trace(2.0 * TS @ TS) + 4.0 * err @ TS @ self.WT @ err
        return -cost_mu, cost_var

    def get_rewards_trajectory(self, states_mu, states_var, actions):
        """:144-149."""
        r, rv = self.get_reward(states_mu[:-1], states_var[:-1], actions)
        rT, rvT = self.get_reward_terminal(states_mu[-1], states_var[-1])
        return torch.cat([r, rT[None]]), torch.cat([rv, rvT[None]])


def action_mpc_to_model(action_mpc, len_horizon, limit_action_change=False,
                        max_change=None, action_prev=None):
    """actions_mappers/normalization_action_mapper.py:21-23 and derivative_action_mapper.py:28-35."""
    a2 = torch.atleast_2d(action_mpc.reshape(len_horizon, -1))
    if not limit_action_change:
        return a2
    mc = _t(max_change)
    a2 = a2 * 2.0 * mc - mc                                       # :30
    a2 = torch.cat([(a2[0] + _t(action_prev))[None], a2[1:]], 0)  # :31
    return Clamp.apply(torch.cumsum(a2, 0), 0.0, 1.0)             # :32-34


def compute_mean_lcb_trajectory(model, reward, actions_mpc, obs_mu, obs_var, len_horizon,
                                iter_ctrl=0, include_time_model=False, limit_action_change=False,
                                max_change=None, action_prev=None, need_grad=True):
    """LCB objective and d/d actions_mpc (controllers/gp_mpc_controller.py:229-285).

    Returns dict(cost, grad (H*Na,), states_mu_pred, states_var_pred, rewards_trajectory,
    rewards_traj_var, cost_traj_mean_lcb) -- the scalar/grad pair plus the five side-effect
    tensors the reference stores on self (:279-283).
    """
    a = _t(actions_mpc).clone().reshape(-1).requires_grad_(need_grad)              # :265-266
    am = action_mpc_to_model(a, len_horizon, limit_action_change, max_change, action_prev)  # :267
    mu, var = predict_trajectory(model, am, _t(obs_mu), _t(obs_var), len_horizon, iter_ctrl,
                                 include_time_model)                               # :268
    r, rv = reward.get_rewards_trajectory(mu, var, am)                             # :269
    ucb = r + reward.kappa * torch.sqrt(rv)                                        # :270
    if reward.clip:
        ucb = Clamp.apply(ucb, float("-inf"), 0.0)                                 # :272-274
    mean_ucb = ucb.mean()
    cost = -mean_ucb                                                               # :275-276
    grad = torch.autograd.grad(cost, a)[0].detach().numpy() if need_grad else None  # :277
    return dict(cost=float(cost.item()), grad=grad, actions_model=am.detach().numpy(),
                states_mu_pred=mu.detach().numpy(), states_var_pred=var.detach().numpy(),
                rewards_trajectory=r.detach().numpy(), rewards_traj_var=rv.detach().numpy(),
                cost_traj_mean_lcb=float(mean_ucb.item()))


def model_from_workload(cfg):
    from oracle.workloads import full_lengthscale
    return OracleModel(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"])


def evaluate_workload(cfg, candidates=None, need_grad=True, model=None):
    """Run the oracle for the given candidate indices of a workload dict; stacks results."""
    model = model or model_from_workload(cfg)
    reward = OracleReward(cfg["reward"])
    idx = range(cfg["B"]) if candidates is None else candidates
    outs = [compute_mean_lcb_trajectory(
        model, reward, cfg["actions"][b], cfg["mu0"], cfg["Sigma0"], cfg["H"], cfg["iter_ctrl"],
        cfg["include_time_model"], cfg["limit_action_change"], cfg["max_change_action_norm"],
        cfg["action_prev"], need_grad) for b in idx]
    res = {k: np.stack([np.asarray(o[k]) for o in outs]) for k in outs[0] if outs[0][k] is not None}
    return res
