"""CPU tests of the batched box-constrained minimisers behind ControllerConfig.batched_candidates:
B independent problems advanced in lock step must each reach what scipy's L-BFGS-B (the reference's optimiser,
gp_mpc_controller.py:133-139) reaches on the same problem."""
import numpy as np
import torch
from scipy.optimize import minimize

from rl_gp_mpc.control_objects.controllers.batched_optim import minimize_box_adam, minimize_box_lbfgs


def quadratic_problem(nb=12, n=10, seed=0):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((nb, n, n))
    A = M @ np.swapaxes(M, 1, 2) / n + 0.3 * np.eye(n)          # SPD, condition number ~ 30
    c = rng.uniform(-0.4, 1.4, size=(nb, n))                    # part of the unconstrained minimisers leave the box
    At, ct = torch.as_tensor(A), torch.as_tensor(c)

    def fun(x):
        r = x - ct
        Ar = torch.einsum("bij,bj->bi", At, r)
        return 0.5 * (r * Ar).sum(1), Ar

    return A, c, fun


def scipy_solution(A, c, x0):
    out = []
    for b in range(len(A)):
        res = minimize(lambda x: (0.5 * (x - c[b]) @ A[b] @ (x - c[b]), A[b] @ (x - c[b])), x0[b], jac=True,
                       method="L-BFGS-B", bounds=[(0, 1)] * A.shape[1], options={"maxiter": 500, "ftol": 1e-15, "gtol": 1e-12})
        out.append(res.fun)
    return np.array(out)


def test_lbfgs_reaches_the_scipy_minimum_on_every_candidate():
    A, c, fun = quadratic_problem()
    x0 = torch.as_tensor(np.random.default_rng(1).uniform(0, 1, size=c.shape))
    want = scipy_solution(A, c, x0.numpy())
    x, f = minimize_box_lbfgs(fun, x0, iters=40)
    assert torch.all((x >= 0) & (x <= 1))
    np.testing.assert_allclose(fun(x)[0].numpy(), f.numpy(), rtol=0, atol=1e-12)   # returned cost belongs to returned x
    np.testing.assert_allclose(f.numpy(), want, rtol=0, atol=1e-6)


def test_lbfgs_needs_far_fewer_evaluations_than_adam():
    A, c, fun = quadratic_problem(seed=3)
    x0 = torch.full(c.shape, 0.5, dtype=torch.float64)
    want = scipy_solution(A, c, x0.numpy())
    _, f15 = minimize_box_lbfgs(fun, x0, iters=15)
    _, fa, _ = minimize_box_adam(fun, x0, iters=30, lr=0.05)
    gap_l = (f15.numpy() - want).max()
    gap_a = (fa.numpy() - want).max()
    assert gap_l < 1e-3 and gap_l < gap_a


def test_candidates_do_not_interact_and_bad_points_are_rejected():
    A, c, fun = quadratic_problem(nb=6, seed=5)
    x0 = torch.as_tensor(np.random.default_rng(2).uniform(0, 1, size=c.shape))
    _, f_all = minimize_box_lbfgs(fun, x0, iters=12)
    for b in (0, 5):
        def fun_b(x, b=b):
            r = x - torch.as_tensor(c[b:b + 1])
            Ar = torch.einsum("bij,bj->bi", torch.as_tensor(A[b:b + 1]), r)
            return 0.5 * (r * Ar).sum(1), Ar
        _, f_one = minimize_box_lbfgs(fun_b, x0[b:b + 1], iters=12)
        np.testing.assert_allclose(f_one.numpy(), f_all[b:b + 1].numpy(), rtol=0, atol=1e-12)

    def fun_nan(x):                       # the objective of candidate 1 is NaN away from its start: it must stay put
        f, g = fun(x)
        moved = (x[1] - x0[1]).abs().max() > 0
        f = f.clone()
        if moved:
            f[1] = float("nan")
        return f, g

    x, f = minimize_box_lbfgs(fun_nan, x0, iters=8)
    assert torch.equal(x[1], x0[1]) and torch.isfinite(f).all()
    assert f[0] < fun(x0)[0][0]


def test_nonconvex_objective_decreases_monotonically():
    def fun(x):                           # shifted Rosenbrock chain inside the box
        z = 4.0 * x - 2.0
        f = (100.0 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1 - z[:, :-1]) ** 2).sum(1)
        g = torch.zeros_like(z)
        g[:, :-1] += -400.0 * z[:, :-1] * (z[:, 1:] - z[:, :-1] ** 2) - 2 * (1 - z[:, :-1])
        g[:, 1:] += 200.0 * (z[:, 1:] - z[:, :-1] ** 2)
        return f, 4.0 * g

    x0 = torch.as_tensor(np.random.default_rng(4).uniform(0.2, 0.8, size=(5, 6)))
    f0 = fun(x0)[0]
    prev = f0
    for iters in (5, 20, 80):
        _, f = minimize_box_lbfgs(fun, x0, iters=iters)
        assert torch.all(f <= prev + 1e-12)
        prev = f
    assert torch.all(prev < 0.05 * f0)


def test_controller_batched_path_with_a_stand_in_engine():
    """Wiring of GpMpcController._get_optimal_actions_batched (no device): both methods drive the engine's batched
    rollout, return the best candidate's actions and store the winner's side effects."""
    import pytest
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.total_config import Config

    H, Na, E = 6, 1, 3
    target = torch.linspace(0.1, 0.9, H * Na, dtype=torch.float64)

    class FakeEngine:
        device = torch.device("cpu")
        calls = 0

        def rollout(self, actions_mpc, obs_mu, obs_var, H_, iter_ctrl=0, limit_action_change=False, max_change=None,
                    action_prev=None, need_grad=True, need_traj=True, out=None):
            FakeEngine.calls += 1
            a = torch.as_tensor(actions_mpc, dtype=torch.float64).reshape(-1, H_ * Na)
            B = a.shape[0]
            o = {"cost": ((a - target) ** 2).sum(1) + 0.25}
            if need_grad:
                o["grad"] = 2 * (a - target)
            if need_traj:
                o.update(states_mu_pred=torch.zeros(B, H_ + 1, E, dtype=torch.float64),
                         states_var_pred=torch.zeros(B, H_ + 1, E, E, dtype=torch.float64),
                         rewards_trajectory=torch.zeros(B, H_ + 1, dtype=torch.float64),
                         rewards_traj_var=torch.zeros(B, H_ + 1, dtype=torch.float64))
            return o

    with pytest.raises(ValueError):
        ControllerConfig(batched_method="newton")
    for method, iters in (("lbfgs", 6), ("adam", 60)):
        cfg = Config(controller_config=ControllerConfig(len_horizon=H, batched_candidates=8, batched_iters=iters,
                                                        batched_method=method))
        c = GpMpcController(-np.ones(E), np.ones(E), -np.ones(Na), np.ones(Na), cfg)
        c.transition_model._engine = FakeEngine()
        c._cost_bound, c._cost_key_values = True, c._cost_fingerprint(cfg.reward)
        FakeEngine.calls = 0
        torch.manual_seed(0)
        act = c._get_optimal_actions_batched(torch.zeros(E, dtype=torch.float64), torch.eye(E, dtype=torch.float64))
        assert FakeEngine.calls == iters + 2                      # iters (+1 value) evaluations + the winner's rollout
        assert c.batched_costs.shape == (8,) and abs(c.last_optim_cost - 0.25) < (1e-10 if method == "lbfgs" else 1e-3)
        assert abs(c.cost_traj_mean_lcb.item() + c.last_optim_cost) < 1e-12
        np.testing.assert_allclose(np.asarray(act).reshape(-1), target.numpy(), atol=1e-5 if method == "lbfgs" else 5e-2)
        np.testing.assert_allclose(c.actions_mpc_previous_iter, target.numpy(), atol=1e-5 if method == "lbfgs" else 5e-2)
