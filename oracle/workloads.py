"""Seeded synthetic workloads for the GP-MPC hot path -- shared by tests and bench.

Recipe: SURVEY.md section 8(d) (x ~ U[0,1]^{N x D}, W ~ N(0,1)^{D x E},
y = 0.05 sin(3 x W) + 1e-3 eps, actions ~ U[0,1], mu0 ~ U[0.25,0.75],
Sigma0 = obs_var * I, hyper-parameters = the reference examples' defaults
(examples/*/config_*.py:41-45).  Pure numpy (PCG64) so the same bits are
produced in the build container and on the GPU box.
"""
import numpy as np

# BASELINE.json configs (C4b = the headline shape, "4-state/2-action")
NAMED = {
    # name: (E, Na, N, H, B, lengthscale, reward preset)
    "C1": dict(E=3, Na=1, N=50, H=15, B=1, ls=0.5, preset="pendulum"),
    "C2": dict(E=3, Na=1, N=200, H=25, B=1024, ls=0.5, preset="pendulum"),
    "C3": dict(E=2, Na=1, N=300, H=40, B=4096, ls=0.5, preset="mountaincar"),
    "C4a": dict(E=2, Na=2, N=500, H=30, B=8192, ls=0.25, preset="process"),
    "C4b": dict(E=4, Na=2, N=500, H=30, B=8192, ls=0.25, preset="generic"),
    "C5": dict(E=8, Na=3, N=1000, H=50, B=65536, ls=0.5, preset="generic"),
}


def _reward_preset(name, E, Na):
    if name == "pendulum":  # examples/pendulum/config_pendulum.py:16-33
        return dict(target_state=[1, 0.5, 0.5], weight_state=[1, 0.1, 0.1], weight_state_terminal=[5, 2, 2],
                    target_action=[0.5], weight_action=[1e-3], exploration_factor=1.0)
    if name == "mountaincar":  # examples/mountain_car/config_mountaincar.py:16-33
        return dict(target_state=[1, 0.5], weight_state=[1, 0], weight_state_terminal=[5, 0],
                    target_action=[0.5], weight_action=[0.05], exploration_factor=1.0)
    if name == "process":  # examples/process_control/config_process_control.py:16-33
        return dict(target_state=[0.5, 0.5], weight_state=[1, 1], weight_state_terminal=[1, 1],
                    target_action=[0, 0], weight_action=[1e-4, 1e-4], exploration_factor=1.0)
    # target 0.5, w_state 1, w_terminal 2, w_action 1e-3, kappa 1
    return dict(target_state=[0.5] * E, weight_state=[1.0] * E, weight_state_terminal=[2.0] * E,
                target_action=[0.5] * Na, weight_action=[1e-3] * Na, exploration_factor=1.0)


def make_workload(name=None, *, E=None, Na=None, N=None, H=None, B=None, ls=0.5, preset="generic",
                  seed=0, noise=1e-5, outputscale=5e-2, obs_var=1e-6, include_time_model=False,
                  distinct_lengthscales=False, limit_action_change=False, use_constraints=False,
                  clip_lower_bound_cost_to_0=False, exploration_factor=None, iter_ctrl=0):
    """Returns a dict of float64 numpy arrays + python scalars describing one workload."""
    if name is not None:
        base = dict(NAMED[name])
        E = E or base["E"]; Na = Na or base["Na"]; N = N or base["N"]; H = H or base["H"]; B = B or base["B"]
        ls = base["ls"]; preset = base["preset"]
        if E != base["E"] or Na != base["Na"]:
            preset = "generic"
    D0 = E + Na
    D = D0 + (1 if include_time_model else 0)
    rng = np.random.default_rng(1000 + seed)
    x = rng.uniform(0.0, 1.0, size=(N, D))
    if include_time_model:
        x[:, -1] = np.arange(N, dtype=np.float64) + iter_ctrl - N  # past control-step indices
    W = rng.standard_normal((D, E))
    xs = x.copy()
    if include_time_model:
        xs[:, -1] = (x[:, -1] - x[:, -1].min()) / max(N, 1)
    y = 0.05 * np.sin(3.0 * xs @ W) + 1e-3 * rng.standard_normal((N, E))
    actions = rng.uniform(0.0, 1.0, size=(B, H, Na))
    mu0 = rng.uniform(0.25, 0.75, size=(E,))
    lengthscale = np.full((E, D0), ls, dtype=np.float64)
    if distinct_lengthscales:  # "trained" hyper-parameters: every GP has its own ARD vector
        lengthscale = lengthscale * rng.uniform(0.8, 1.6, size=(E, D0))
    outputscale_v = np.full((E,), outputscale)
    noise_v = np.full((E,), noise)
    if distinct_lengthscales:
        outputscale_v = outputscale_v * rng.uniform(0.7, 1.4, size=(E,))
        noise_v = noise_v * rng.uniform(0.7, 1.4, size=(E,))
    reward = _reward_preset(preset, E, Na)
    if exploration_factor is not None:
        reward["exploration_factor"] = float(exploration_factor)
    reward.update(use_constraints=use_constraints, clip_lower_bound_cost_to_0=clip_lower_bound_cost_to_0,
                  state_min=list(np.linspace(0.1, 0.3, E)), state_max=list(np.linspace(0.9, 0.8, E)))
    return dict(
        name=name or "custom", seed=seed, E=E, Na=Na, D=D, N=N, H=H, B=B,
        x=x, y=y, actions=actions, mu0=mu0, obs_var=obs_var, Sigma0=obs_var * np.eye(E),
        lengthscale=lengthscale, lengthscale_time=100.0, outputscale=outputscale_v, noise=noise_v,
        include_time_model=include_time_model, iter_ctrl=iter_ctrl,
        limit_action_change=limit_action_change,
        max_change_action_norm=list(np.linspace(0.1, 0.2, Na)),
        action_prev=list(rng.uniform(0.3, 0.7, size=(Na,))),
        reward=reward,
    )


def full_lengthscale(cfg):
    """(E, D) lengthscale matrix incl. the time column (functions_process_config.py:18-28)."""
    ls = np.asarray(cfg["lengthscale"], dtype=np.float64)
    if cfg["include_time_model"]:
        ls = np.concatenate([ls, np.full((ls.shape[0], 1), cfg["lengthscale_time"])], axis=1)
    return ls


def algorithmic_bytes_per_prediction(E, D, N, elem_bytes=8):
    """B_alg = s * (E N^2 + E N + N D)."""
    return elem_bytes * (E * N * N + E * N + N * D)


def algorithmic_flops_per_prediction(E, D, N):
    """F_alg = 2 [P N^2 (D+3) + E N^2 + E N (D^2+3D+3) + P N D^2]."""
    P = E * (E + 1) // 2
    return 2 * (P * N * N * (D + 3) + E * N * N + E * N * (D * D + 3 * D + 3) + P * N * D * D)
