"""GPU: hyper-parameter objective on the device: exact-GP log marginal likelihood + gradient,
and the reference-style training procedure that drives it (gp_model.py:193-306)."""
import queue as queue_mod

import numpy as np
import pytest
import torch

from oracle import gpmpc_oracle as orc
from oracle.workloads import full_lengthscale, make_workload

pytestmark = pytest.mark.gpu


def cpu_lml(x, y, ls, s2, noise):
    """float64 torch reference: sum over GPs is NOT taken; returns (E,) LML and autograd gradients."""
    x = torch.as_tensor(x); y = torch.as_tensor(y)
    ls = torch.as_tensor(ls).clone().requires_grad_(True)
    s2 = torch.as_tensor(s2).clone().requires_grad_(True)
    noise = torch.as_tensor(noise).clone().requires_grad_(True)
    K = orc.gram_matrix(x, ls, s2) + noise[:, None, None] * torch.eye(x.shape[0], dtype=torch.float64)
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve(y.t()[:, :, None], L)[:, :, 0]
    lml = -0.5 * (y.t() * alpha).sum(1) - torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(1) \
        - 0.5 * x.shape[0] * np.log(2 * np.pi)
    grads = []
    for a in range(lml.shape[0]):
        g = torch.autograd.grad(lml[a], (ls, s2, noise), retain_graph=True)
        grads.append((g[0][a].numpy(), g[1][a].item(), g[2][a].item()))
    return lml.detach().numpy(), grads


@pytest.mark.parametrize("kw", [dict(E=2, Na=1, N=70, ls=0.5, seed=51, noise=1e-3),
                                dict(E=3, Na=2, N=150, ls=0.4, seed=52, noise=1e-4, distinct_lengthscales=True)])
def test_marginal_likelihood_and_gradient_match_autograd(kw):
    from rl_gp_mpc import _cabi
    cfg = make_workload(H=2, B=1, **kw)
    ls = full_lengthscale(cfg)
    eng = _cabi.Engine()
    eng.prepare(cfg["x"], cfg["y"], ls, cfg["outputscale"], cfg["noise"])
    out = eng.mll(cfg["y"]).cpu().numpy()
    lml, grads = cpu_lml(cfg["x"], cfg["y"], ls, cfg["outputscale"], cfg["noise"])
    D = cfg["D"]
    for a in range(cfg["E"]):
        assert abs(out[a, 0] - lml[a]) <= 1e-8 * max(1.0, abs(lml[a]))
        scale = max(1.0, np.abs(grads[a][0]).max(), abs(grads[a][1]), abs(grads[a][2]))
        np.testing.assert_allclose(out[a, 3:3 + D], grads[a][0], rtol=0, atol=1e-7 * scale)
        assert abs(out[a, 1] - grads[a][1]) <= 1e-7 * scale
        assert abs(out[a, 2] - grads[a][2]) <= 1e-7 * scale


def test_fit_eval_graph_equals_prepare_plus_mll_and_isolates_a_failing_gp():
    """gpmpc_fit_eval (one CUDA-graph launch per objective evaluation of the fit) returns what gpmpc_prepare + gpmpc_mll
    return, also when replayed with other hyper-parameters, leaves the handle prepared, and reports a GP whose kernel
    matrix is not positive definite through `info` without touching the other GPs' rows."""
    from rl_gp_mpc import _cabi
    cfg = make_workload(E=3, Na=2, N=150, H=2, B=1, ls=0.4, seed=53, noise=1e-4, distinct_lengthscales=True)
    ls = torch.as_tensor(full_lengthscale(cfg))
    s2 = torch.as_tensor(cfg["outputscale"]).reshape(-1)
    nz = torch.as_tensor(cfg["noise"]).reshape(-1)
    E, D = ls.shape
    eng = _cabi.Engine()
    x_dev = torch.as_tensor(cfg["x"], dtype=torch.float64).to(eng.device).contiguous()
    y_dev = torch.as_tensor(cfg["y"], dtype=torch.float64).to(eng.device).contiguous()
    ref = _cabi.Engine()
    for scale in (1.0, 1.3, 0.8):                       # capture, then two replays
        theta = torch.cat([ls * scale, (s2 * scale)[:, None], nz[:, None]], dim=1)
        out, info = eng.fit_eval(x_dev, y_dev, theta)
        assert int(info.abs().sum()) == 0
        ref.prepare(cfg["x"], cfg["y"], ls * scale, s2 * scale, nz)
        want = ref.mll(cfg["y"]).cpu()
        np.testing.assert_allclose(out.numpy(), want.numpy(), rtol=1e-12, atol=1e-9)
    # a larger prepare() on the same handle re-allocates its buffers: the captured graph must not be replayed
    big = make_workload(E=3, Na=2, N=400, H=2, B=1, ls=0.4, seed=54, noise=1e-4)
    eng.prepare(big["x"], big["y"], full_lengthscale(big), big["outputscale"], big["noise"])
    out2, info2 = eng.fit_eval(x_dev, y_dev, theta)
    assert int(info2.abs().sum()) == 0
    np.testing.assert_allclose(out2.numpy(), want.numpy(), rtol=1e-12, atol=1e-9)
    iK, beta = eng.factorization()                      # the handle is prepared at the last theta
    iK_ref, beta_ref = ref.factorization()
    np.testing.assert_allclose(beta.cpu().numpy(), beta_ref.cpu().numpy(), rtol=1e-12, atol=1e-12)
    # GP 1 with a huge lengthscale and no noise: numerically singular K -> info[1] != 0, rows 0 and 2 as before
    theta_bad = theta.clone()
    theta_bad[1, :D] = 1e6
    theta_bad[1, D + 1] = 0.0
    out_bad, info_bad = eng.fit_eval(x_dev, y_dev, theta_bad)
    assert int(info_bad[1]) != 0 and int(info_bad[0]) == 0 and int(info_bad[2]) == 0
    np.testing.assert_allclose(out_bad[[0, 2]].numpy(), out[[0, 2]].numpy(), rtol=1e-12, atol=1e-9)


def test_training_procedure_improves_the_marginal_likelihood_and_feeds_the_controller():
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.config_classes.observation_config import ObservationConfig
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.config_classes.total_config import Config
    from rl_gp_mpc.config_classes.training_config import TrainingConfig
    from rl_gp_mpc.control_objects.models.gp_model import GpStateTransitionModel
    cfg = make_workload(E=2, Na=1, N=80, H=4, B=1, ls=0.5, seed=53, preset="mountaincar")
    r = cfg["reward"]
    config = Config(
        observation_config=ObservationConfig(obs_var_norm=[1e-6] * 2),
        reward_config=RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                                   weight_state_terminal=list(r["weight_state_terminal"]),
                                   target_action_norm=list(r["target_action"]), weight_action=list(r["weight_action"])),
        controller_config=ControllerConfig(len_horizon=4),
        training_config=TrainingConfig(lr_train=0.5, iter_train=8, training_frequency=1000),
        model_config=ModelConfig(gp_init={"noise_covar.noise": [1e-3, 1e-3], "base_kernel.lengthscale": [2.0, 2.0],
                                          "outputscale": [0.5, 0.5]},
                                 min_std_noise=1e-3, max_std_noise=1e-1, min_outputscale=1e-4, max_outputscale=1.0,
                                 min_lengthscale=5e-2, max_lengthscale=10.0))
    ctrl = GpMpcController(-np.ones(2), np.ones(2), -np.ones(1), np.ones(1), config)
    ctrl.memory.model_inputs[:cfg["N"]] = torch.as_tensor(cfg["x"])
    ctrl.memory.model_targets[:cfg["N"]] = torch.as_tensor(cfg["y"])
    ctrl.memory.len_mem_model = cfg["N"]
    ctrl.get_action(np.zeros(2))
    tm = ctrl.transition_model
    assert tm.engine.uses_uniform_path()
    # direct call of the reference-signature static method
    torch.manual_seed(3)
    q = queue_mod.Queue()
    st = tm.save_state(); st.to_arrays()
    GpStateTransitionModel.train(q, st, 0.5, 8, 1e-3)
    new = q.get(timeout=5)
    assert len(new) == 2 and set(new[0]) == {"covar_module.base_kernel.lengthscale", "covar_module.outputscale", "likelihood.noise"}
    old_ls = np.full((2, 3), 2.0)
    eng = tm.engine
    def neg_mll(ls, s2, nz):
        eng.prepare(cfg["x"], cfg["y"], ls, s2, nz)
        return -eng.mll(cfg["y"])[:, 0].cpu().numpy() / cfg["N"]
    before = neg_mll(old_ls, [0.5, 0.5], [1e-3, 1e-3])
    after = neg_mll(np.stack([p["covar_module.base_kernel.lengthscale"][0] for p in new]),
                    [float(p["covar_module.outputscale"]) for p in new], [float(p["likelihood.noise"][0]) for p in new])
    assert np.all(after <= before + 1e-12) and np.any(after < before - 1e-3)
    for p in new:
        assert np.all(p["covar_module.base_kernel.lengthscale"] >= 5e-2) and 1e-6 <= p["likelihood.noise"][0] <= 1e-2
    # asynchronous flow through the controller (gp_mpc_controller.py:201-227)
    ctrl.start_training_process()
    ctrl.p_train.join(timeout=120)
    assert not ctrl.p_train.is_alive()
    ctrl.check_and_close_processes()
    assert ctrl.p_train._closed
    act = ctrl.get_action(np.zeros(2))                       # per-GP hyper-parameters now -> general kernel path
    assert act.shape == (1,) and np.isfinite(ctrl.last_optim_cost)


def test_lockstep_fit_of_all_gps_equals_the_serial_fit():
    """GpStateTransitionModel.train(lockstep=True) batches the objective evaluations of the E concurrent LBFGS fits into
    one gpmpc_prepare + gpmpc_mll per round; every GP must end where the one-GP-after-the-other procedure ends
    (same random restarts: same torch seed)."""
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.control_objects.models.gp_model import GpStateTransitionModel
    cfg = make_workload(E=3, Na=1, N=120, H=3, B=1, ls=0.5, seed=57)
    mc = ModelConfig(gp_init={"noise_covar.noise": [1e-3] * 3, "base_kernel.lengthscale": [1.5] * 3, "outputscale": [0.3] * 3},
                     min_std_noise=1e-3, max_std_noise=1e-1, min_outputscale=1e-4, max_outputscale=1.0,
                     min_lengthscale=5e-2, max_lengthscale=10.0)
    tm = GpStateTransitionModel(mc, dim_state=3, dim_action=1)
    tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    out = {}
    for lockstep in (False, True):
        torch.manual_seed(5)
        q = queue_mod.Queue()
        st = tm.save_state(); st.to_arrays()
        GpStateTransitionModel.train(q, st, 0.5, 6, 1e-3, lockstep=lockstep)
        out[lockstep] = q.get(timeout=5)
    eng = tm.engine
    def neg_mll(params):
        eng.prepare(cfg["x"], cfg["y"], np.stack([p["covar_module.base_kernel.lengthscale"][0] for p in params]),
                    [float(p["covar_module.outputscale"]) for p in params], [float(p["likelihood.noise"][0]) for p in params])
        return -eng.mll(cfg["y"])[:, 0].cpu().numpy() / cfg["N"]
    np.testing.assert_allclose(neg_mll(out[True]), neg_mll(out[False]), rtol=0, atol=1e-6)
    for a, b in zip(out[True], out[False]):
        np.testing.assert_allclose(a["covar_module.base_kernel.lengthscale"], b["covar_module.base_kernel.lengthscale"], rtol=1e-5)


def test_fit_ends_where_the_torch_lbfgs_procedure_of_the_reference_ends():
    """train() drives its own generator form of LBFGS / strong Wolfe; the reference drives torch.optim.LBFGS
    (gp_model.py:262-277).  Same random restart (torch seed), same objective (device marginal likelihood): the final
    losses must agree within 1e-6."""
    from rl_gp_mpc import _cabi
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.control_objects.models.gp_model import GpStateTransitionModel
    cfg = make_workload(E=2, Na=1, N=100, H=3, B=1, ls=0.5, seed=58)
    E, D, n = 2, 3, cfg["N"]
    mc = ModelConfig(gp_init={"noise_covar.noise": [1e-3] * E, "base_kernel.lengthscale": [1.5] * E, "outputscale": [0.3] * E},
                     min_std_noise=1e-3, max_std_noise=1e-1, min_outputscale=1e-4, max_outputscale=1.0,
                     min_lengthscale=5e-2, max_lengthscale=10.0)
    tm = GpStateTransitionModel(mc, dim_state=E, dim_action=1)
    tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    lr, iters = 0.5, 5
    torch.manual_seed(9)
    q = queue_mod.Queue()
    st = tm.save_state(); st.to_arrays()
    GpStateTransitionModel.train(q, st, lr, iters, 1e-3)
    new = q.get(timeout=5)
    eng = _cabi.Engine()
    x = torch.as_tensor(cfg["x"]); y = torch.as_tensor(cfg["y"])

    def neg_mll_one(idx, theta):
        eng.prepare(x, y[:, idx:idx + 1].contiguous(), theta[:D].reshape(1, D), theta[D:D + 1], theta[D + 1:D + 2])
        out = eng.mll(y[:, idx:idx + 1].contiguous())[0].cpu()
        return -out[0] / n, -torch.cat([out[3:3 + D], out[1:3]]) / n

    torch.manual_seed(9)
    starts = [torch.rand(D + 2, dtype=torch.float64) for _ in range(E)]       # as train() draws them
    lo = torch.tensor([5e-2] * D + [1e-4, 1e-6], dtype=torch.float64)
    hi = torch.tensor([10.0] * D + [1.0, 1e-2], dtype=torch.float64)
    for idx in range(E):
        class _F(torch.autograd.Function):
            @staticmethod
            def forward(ctx, theta):
                loss, grad = neg_mll_one(idx, theta.detach())
                ctx.save_for_backward(grad)
                return loss

            @staticmethod
            def backward(ctx, g):
                return g * ctx.saved_tensors[0]
        prev = torch.tensor([1.5] * D + [0.3, 1e-3], dtype=torch.float64)
        best = float(_F.apply(prev))
        p0 = ((lo + starts[idx] * (hi - lo) - lo) / (hi - lo)).clamp(1e-6, 1 - 1e-6)
        raw = (torch.log(p0) - torch.log1p(-p0)).requires_grad_(True)
        opt = torch.optim.LBFGS([raw], lr=lr, line_search_fn="strong_wolfe")
        best_theta = prev
        for _ in range(iters):
            def closure():
                opt.zero_grad()
                loss = _F.apply(lo + (hi - lo) * torch.sigmoid(raw))
                loss.backward()
                return loss
            loss = float(opt.step(closure))
            if loss < best:
                best, best_theta = loss, (lo + (hi - lo) * torch.sigmoid(raw)).detach().clone()
        got = torch.cat([torch.as_tensor(new[idx]["covar_module.base_kernel.lengthscale"]).reshape(-1),
                         torch.as_tensor(new[idx]["covar_module.outputscale"]).reshape(1),
                         torch.as_tensor(new[idx]["likelihood.noise"]).reshape(1)])
        assert abs(float(neg_mll_one(idx, got)[0]) - float(neg_mll_one(idx, best_theta)[0])) <= 1e-6
        np.testing.assert_allclose(got.numpy(), best_theta.numpy(), rtol=1e-3, atol=1e-9)   # (flat directions: the loss is the criterion)
