"""GPU parity tests proper: the CUDA path through the C ABI vs (a) the golden vectors produced by the
reference's own code, (b) the CPU oracle on seeded workloads, (c) size-independent properties at full size.

Tolerance (float64 path, stated per north_star): absolute 1e-8 on costs / states / variances / rewards and
1e-7 on gradients and V, for training sets with cond(K + noise I) up to ~1e7 (the reference's hyper-parameter
regime, noise 1e-5).  Measured errors on B200 are 1e-11 or better (smoke: profiles/r01_s8_smoke.txt)."""
import numpy as np
import pytest
import torch

from oracle import gpmpc_oracle as orc
from oracle.workloads import full_lengthscale, make_workload
from tests.golden_utils import big_case_names, case_names, load_big_case, load_case

pytestmark = pytest.mark.gpu

ATOL = 1e-8
ATOL_GRAD = 1e-7


def make_engine(cfg, path=0):
    """path 0: automatic kernel selection (uniform-kernel fast path when all GPs share their hyper-parameters),
    path 1: force the general per-pair path."""
    from rl_gp_mpc import _cabi
    eng = _cabi.Engine()
    eng.set_path(path)
    eng.prepare(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"])
    r = cfg["reward"]
    W = np.diag(np.concatenate([r["weight_state"], r["weight_action"]]).astype(float))
    WT = np.diag(np.asarray(r["weight_state_terminal"], float))
    tgt = np.concatenate([r["target_state"], r["target_action"]]).astype(float)
    eng.set_cost(tgt, W, WT, r["exploration_factor"], r["use_constraints"], r["state_min"], r["state_max"],
                 r["clip_lower_bound_cost_to_0"])
    return eng


def rollout(eng, cfg, actions=None, need_grad=True):
    a = cfg["actions"] if actions is None else actions
    out = eng.rollout(a, cfg["mu0"], cfg["Sigma0"], cfg["H"], cfg["iter_ctrl"], cfg["limit_action_change"],
                      cfg["max_change_action_norm"], cfg["action_prev"], need_grad=need_grad)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("name", case_names())
def test_cuda_matches_reference_golden(name, path):
    cfg, gold = load_case(name)
    eng = make_engine(cfg, path)
    assert eng.uses_uniform_path() == (path == 0 and name != "distinct_ls_e4")
    iK, beta = eng.factorization()
    scale = np.abs(gold["iK"]).max()
    assert np.abs(iK.cpu().numpy() - gold["iK"]).max() <= 1e-9 * scale          # calculate_factorizations
    assert np.abs(beta.cpu().numpy() - gold["beta"]).max() <= 1e-9 * max(1.0, np.abs(gold["beta"]).max())
    E = cfg["E"]
    M, S, V = eng.predict_step(gold["step_in_mu"][None], gold["step_in_var"][None, :E, :E])   # predict_next_state_change
    np.testing.assert_allclose(M.cpu().numpy()[0], gold["step_M"][0], rtol=0, atol=ATOL)
    np.testing.assert_allclose(S.cpu().numpy()[0], gold["step_S"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(V.cpu().numpy()[0], gold["step_V"], rtol=0, atol=ATOL_GRAD)
    out = rollout(eng, cfg)                                                        # compute_mean_lcb_trajectory
    np.testing.assert_allclose(out["cost"], gold["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(out["grad"], gold["grad"], rtol=0, atol=ATOL_GRAD)
    np.testing.assert_allclose(out["states_mu_pred"], gold["states_mu_pred"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(out["states_var_pred"], gold["states_var_pred"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(out["rewards_trajectory"], gold["rewards_trajectory"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(out["rewards_traj_var"], gold["rewards_traj_var"], rtol=0, atol=ATOL)
    fwd = rollout(eng, cfg, need_grad=False)                                       # value-only kernel variant
    np.testing.assert_allclose(fwd["cost"], gold["cost"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("name", big_case_names())
def test_cuda_matches_reference_golden_ill_conditioned(name, path):
    """Vectors of the verbatim reference at N = 200 (H = 25) and N = 500, noise 1e-5: cond(K + noise I) ~ 1e6 .. 1e7, the
    regime in which beta^T L beta and tr(iK L) cancel by ~1e8 (gp_model.py:169-176).  The inverse itself is compared
    through its diagonal and row sums (the golden file does not carry the 4 MB matrix).  Horizon 25: tolerance x10 --
    the CPU oracle and the reference, from bit-identical iK / beta, already differ by 2e-8 there
    (tests/test_oracle_vs_golden.py)."""
    cfg, gold = load_big_case(name)
    if path == 1 and name.endswith("distinct"):
        pytest.skip("already on the general path")
    eng = make_engine(cfg, path)
    iK, beta = eng.factorization()
    iK = iK.cpu().numpy()
    scale = float(gold["iK_absmax"])
    assert np.abs(np.diagonal(iK, axis1=1, axis2=2) - gold["iK_diag"]).max() <= 1e-7 * scale
    assert np.abs(iK.sum(axis=2) - gold["iK_rowsum"]).max() <= 1e-6 * scale
    assert np.abs(beta.cpu().numpy() - gold["beta"]).max() <= 1e-7 * max(1.0, np.abs(gold["beta"]).max())
    E = cfg["E"]
    M, S, V = eng.predict_step(gold["step_in_mu"][None], gold["step_in_var"][None, :E, :E])
    np.testing.assert_allclose(M.cpu().numpy()[0], gold["step_M"][0], rtol=0, atol=ATOL)
    np.testing.assert_allclose(S.cpu().numpy()[0], gold["step_S"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(V.cpu().numpy()[0], gold["step_V"], rtol=0, atol=ATOL_GRAD)
    f = 10.0 if cfg["H"] > 10 else 1.0
    out = rollout(eng, cfg)
    np.testing.assert_allclose(out["cost"], gold["cost"], rtol=0, atol=ATOL * f)
    np.testing.assert_allclose(out["grad"], gold["grad"], rtol=0, atol=ATOL_GRAD * f)
    np.testing.assert_allclose(out["states_mu_pred"], gold["states_mu_pred"], rtol=0, atol=ATOL * f)
    np.testing.assert_allclose(out["states_var_pred"], gold["states_var_pred"], rtol=0, atol=ATOL * f)
    np.testing.assert_allclose(out["rewards_trajectory"], gold["rewards_trajectory"], rtol=0, atol=ATOL * f)
    np.testing.assert_allclose(out["rewards_traj_var"], gold["rewards_traj_var"], rtol=0, atol=ATOL * f * 10)


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("kw", [
    dict(name="C2", B=1, seed=51),        # E=3, N=200, H=25
    dict(name="C3", B=1, seed=52),        # E=2, N=300, H=40
    dict(name="C4b", B=1, seed=53),       # E=4, N=500, H=30: the headline shape, one candidate
])
def test_full_horizon_of_the_baseline_shapes_matches_cpu_oracle(kw, path):
    """One candidate of BASELINE.json configs 2, 3 and 4 at their FULL horizon (25 / 40 / 30 steps) and training-set
    size, both kernel paths, against the CPU oracle.  Long-horizon tolerance: 10x the short-horizon one (see
    test_cuda_matches_reference_golden_ill_conditioned)."""
    cfg = make_workload(**kw)
    eng = make_engine(cfg, path)
    want = orc.evaluate_workload(cfg)
    got = rollout(eng, cfg)
    np.testing.assert_allclose(got["cost"], want["cost"], rtol=0, atol=ATOL * 10)
    np.testing.assert_allclose(got["grad"], want["grad"], rtol=0, atol=ATOL_GRAD * 10)
    np.testing.assert_allclose(got["states_mu_pred"], want["states_mu_pred"], rtol=0, atol=ATOL * 10)
    np.testing.assert_allclose(got["states_var_pred"], want["states_var_pred"], rtol=0, atol=ATOL * 10)
    np.testing.assert_allclose(got["rewards_traj_var"], want["rewards_traj_var"], rtol=0, atol=ATOL * 100)


@pytest.mark.parametrize("path", [0, 1])
def test_c5_dims_full_training_set_one_step(path):
    """BASELINE.json config 5 dims (E=8, Na=3, D=11) at the FULL training-set size N=1000: one moment-matching step
    (objective only, H=1) against the CPU oracle on both kernel paths."""
    cfg = make_workload("C5", B=2, H=1, seed=54)
    eng = make_engine(cfg, path)
    want = orc.evaluate_workload(cfg, need_grad=False)
    got = rollout(eng, cfg, need_grad=False)
    np.testing.assert_allclose(got["cost"], want["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["states_mu_pred"], want["states_mu_pred"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["states_var_pred"], want["states_var_pred"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("kw", [
    dict(name="C2", B=3, H=6, seed=11),                                   # N=200 (pendulum dims)
    dict(name="C3", B=2, H=5, seed=12),                                   # N=300 (mountain car dims)
    dict(name="C4a", B=2, H=3, seed=13),                                  # N=500, E=2
    dict(name="C4b", B=2, H=3, seed=14, N=320),                           # headline dims, N not a multiple of 64
    dict(E=5, Na=2, N=130, H=3, B=2, ls=0.6, seed=15, distinct_lengthscales=True),
    dict(E=8, Na=3, N=96, H=2, B=2, ls=0.7, seed=16),                     # C5 dims, small N
    dict(E=3, Na=1, N=70, H=4, B=3, ls=0.5, seed=17, limit_action_change=True, include_time_model=True, iter_ctrl=9),
    # tensor-core sweeps of the uniform path (E >= 6) with zero-padded k steps / unused output columns, N off the 32-row grid
    dict(E=6, Na=2, N=150, H=3, B=2, ls=0.7, seed=18),
    dict(E=7, Na=1, N=100, H=2, B=3, ls=0.8, seed=19, include_time_model=True),
])
@pytest.mark.parametrize("path", [0, 1])
def test_cuda_matches_cpu_oracle(kw, path):
    cfg = make_workload(**kw)
    if path == 1 and kw.get("distinct_lengthscales"):
        pytest.skip("already on the general path")
    eng = make_engine(cfg, path)
    want = orc.evaluate_workload(cfg)
    got = rollout(eng, cfg)
    np.testing.assert_allclose(got["cost"], want["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["grad"], want["grad"], rtol=0, atol=ATOL_GRAD)
    np.testing.assert_allclose(got["states_mu_pred"], want["states_mu_pred"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["states_var_pred"], want["states_var_pred"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["rewards_traj_var"], want["rewards_traj_var"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("path", [0, 1])
def test_far_rows_with_wide_input_distribution(path):
    """Short lengthscale + very wide state distribution: some training points have a row exponent kap_i < -600 whose
    row factor exp(kap_i) is taken out of the sweep; the uniform kernels then carry the residual shift per element
    (uni_fwd_cols<SH=true>), and those rows still contribute (kap_i + kap_j + cross term ~ 0 for neighbours)."""
    cfg = make_workload(E=2, Na=1, N=150, H=3, B=2, ls=0.02, seed=31, obs_var=4.0)
    eng = make_engine(cfg, path)
    want = orc.evaluate_workload(cfg)
    got = rollout(eng, cfg)
    np.testing.assert_allclose(got["cost"], want["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["grad"], want["grad"], rtol=0, atol=ATOL_GRAD)
    np.testing.assert_allclose(got["states_mu_pred"], want["states_mu_pred"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["states_var_pred"], want["states_var_pred"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("distinct", [False, True])
def test_batched_equals_looped_and_is_order_independent(distinct):
    """B candidates in one call == B single-candidate calls: no cross-candidate arithmetic.  Not bit-for-bit:
    the N^2 partial sums are combined with shared-memory float64 atomics in warp-scheduling order, and the
    covariance sums cancel by ~1e8, so repeated evaluations of the SAME candidate differ by ~1e-10."""
    cfg = make_workload("C2", B=300, H=4, seed=21, distinct_lengthscales=distinct)   # > 148 CTAs: persistent loop
    eng = make_engine(cfg)
    full = rollout(eng, cfg)
    for b in (0, 147, 148, 299):
        one = rollout(eng, cfg, actions=cfg["actions"][b:b + 1])
        np.testing.assert_allclose(one["cost"][0], full["cost"][b], rtol=0, atol=1e-9)
        np.testing.assert_allclose(one["grad"][0], full["grad"][b], rtol=0, atol=1e-8)
        np.testing.assert_allclose(one["states_var_pred"][0], full["states_var_pred"][b], rtol=0, atol=1e-9)
    perm = np.random.default_rng(0).permutation(300)
    shuf = rollout(eng, cfg, actions=cfg["actions"][perm])
    np.testing.assert_allclose(shuf["cost"], full["cost"][perm], rtol=0, atol=1e-9)


def test_gradient_against_central_differences():
    cfg = make_workload(E=2, Na=2, N=90, H=4, B=1, ls=0.3, preset="process", seed=22, noise=1e-4)
    eng = make_engine(cfg)
    base = rollout(eng, cfg)
    a0 = cfg["actions"][0].reshape(-1)
    eps = 1e-5
    pert = np.stack([a0 + eps * s * np.eye(a0.size)[k] for k in range(a0.size) for s in (1, -1)])
    c = rollout(eng, cfg, actions=pert.reshape(-1, cfg["H"], cfg["Na"]), need_grad=False)["cost"]
    fd = (c[0::2] - c[1::2]) / (2 * eps)
    np.testing.assert_allclose(base["grad"][0], fd, rtol=0, atol=5e-7)


@pytest.mark.parametrize("distinct", [False, True])
def test_full_size_properties_headline_shape(distinct):
    """BASELINE config 4 sizes (N=500, E=4, Na=2) on a slice of the batch: invariants that need no oracle."""
    cfg = make_workload("C4b", B=64, H=5, seed=23, distinct_lengthscales=distinct)
    eng = make_engine(cfg)
    out = rollout(eng, cfg)
    var = out["states_var_pred"]
    assert np.all(np.isfinite(out["cost"])) and np.all(np.isfinite(out["grad"]))
    assert np.abs(var - np.swapaxes(var, -1, -2)).max() <= 1e-15                 # covariances stay symmetric
    assert np.linalg.eigvalsh(var).min() > 0                                    # ... and positive definite
    assert np.all(out["rewards_traj_var"] >= 0)
    # LCB definition (gp_mpc_controller.py:270-276) recomputed on the host from the returned trajectories
    ucb = out["rewards_trajectory"] + cfg["reward"]["exploration_factor"] * np.sqrt(out["rewards_traj_var"])
    np.testing.assert_allclose(out["cost"], -ucb.mean(1), rtol=0, atol=1e-14)
    # duplicated candidates give identical results; two oracle spot checks at full N
    dup = rollout(eng, cfg, actions=np.concatenate([cfg["actions"][:2], cfg["actions"][:2]]))
    np.testing.assert_allclose(dup["cost"][:2], dup["cost"][2:], rtol=0, atol=1e-9)
    want = orc.evaluate_workload(cfg, candidates=[5])
    np.testing.assert_allclose(out["cost"][5], want["cost"][0], rtol=0, atol=ATOL)
    np.testing.assert_allclose(out["grad"][5], want["grad"][0], rtol=0, atol=ATOL_GRAD)


def test_per_candidate_initial_state_and_nan_propagation():
    cfg = make_workload("C1", B=4, H=3, seed=24)
    eng = make_engine(cfg)
    mu = np.stack([cfg["mu0"] + 0.01 * b for b in range(4)])
    var = np.stack([cfg["Sigma0"] * (1 + b) for b in range(4)])
    out = eng.rollout(cfg["actions"], mu, var, cfg["H"])
    torch.cuda.synchronize()
    model = orc.model_from_workload(cfg)
    rew = orc.OracleReward(cfg["reward"])
    for b in (0, 3):
        ref = orc.compute_mean_lcb_trajectory(model, rew, cfg["actions"][b], mu[b], var[b], cfg["H"])
        assert abs(out["cost"][b].item() - ref["cost"]) < ATOL
        np.testing.assert_allclose(out["grad"][b].cpu().numpy(), ref["grad"], rtol=0, atol=ATOL_GRAD)
    mu[1, 0] = np.nan                                   # NaN flows out like in the reference, others unaffected
    out2 = eng.rollout(cfg["actions"], mu, var, cfg["H"], need_grad=False)
    torch.cuda.synchronize()
    assert np.isnan(out2["cost"][1].item()) and abs(out2["cost"][0].item() - out["cost"][0].item()) < 1e-9


def test_error_codes():
    from rl_gp_mpc import _cabi
    eng = _cabi.Engine()
    with pytest.raises(_cabi.GpmpcError, match="prepare"):
        eng.predict_step(np.zeros((1, 3)), np.eye(2)[None])
    cfg = make_workload("C1", B=1, H=3)
    with pytest.raises(_cabi.GpmpcError, match="positive definite"):       # K - 2 s2 I is negative definite
        eng.prepare(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], -2.0 * cfg["outputscale"])
    eng.prepare(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"])
    with pytest.raises(_cabi.GpmpcError, match="set_cost"):
        eng.Na = 1
        eng.rollout(cfg["actions"], cfg["mu0"], cfg["Sigma0"], 3)


def test_controller_api_drop_in():
    """The reference-facing call path: GpMpcController.compute_mean_lcb_trajectory and the model methods."""
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.actions_config import ActionsConfig
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.config_classes.observation_config import ObservationConfig
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.config_classes.total_config import Config
    cfg, gold = load_case("pendulum_c1")
    r = cfg["reward"]
    config = Config(
        observation_config=ObservationConfig(obs_var_norm=[cfg["obs_var"]] * 3),
        reward_config=RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                                   weight_state_terminal=list(r["weight_state_terminal"]),
                                   target_action_norm=list(r["target_action"]), weight_action=list(r["weight_action"]),
                                   exploration_factor=r["exploration_factor"]),
        actions_config=ActionsConfig(), controller_config=ControllerConfig(len_horizon=cfg["H"]),
        model_config=ModelConfig(gp_init={"noise_covar.noise": list(cfg["noise"]),
                                          "base_kernel.lengthscale": [list(v) for v in cfg["lengthscale"]],
                                          "outputscale": list(cfg["outputscale"])},
                                 min_std_noise=1e-4, max_std_noise=1.0, min_outputscale=1e-6, max_outputscale=10.0,
                                 min_lengthscale=1e-3, max_lengthscale=1e3))
    ctrl = GpMpcController(-np.ones(3), np.ones(3), -np.ones(1), np.ones(1), config)
    tm = ctrl.transition_model
    tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    assert np.abs(tm.iK.numpy() - gold["iK"]).max() <= 1e-9 * np.abs(gold["iK"]).max()
    M, S, V = tm.predict_next_state_change(torch.as_tensor(gold["step_in_mu"]), torch.as_tensor(gold["step_in_var"]))
    assert M.shape == (1, 3) and S.shape == (3, 3) and V.shape == (4, 3)                   # gp_model.py:180
    np.testing.assert_allclose(S.numpy(), gold["step_S"], rtol=0, atol=ATOL)
    mu_t, var_t = tm.predict_trajectory(torch.as_tensor(cfg["actions"][0]), torch.as_tensor(cfg["mu0"]),
                                        torch.as_tensor(cfg["Sigma0"]), cfg["H"], 0)
    np.testing.assert_allclose(mu_t.numpy(), gold["states_mu_pred"][0], rtol=0, atol=ATOL)
    np.testing.assert_allclose(var_t.numpy(), gold["states_var_pred"][0], rtol=0, atol=ATOL)
    obs_mu, obs_var = torch.as_tensor(cfg["mu0"]), torch.as_tensor(cfg["Sigma0"])
    c, g = ctrl.compute_mean_lcb_trajectory(cfg["actions"][1].reshape(-1), obs_mu, obs_var)
    assert isinstance(c, float) and isinstance(g, np.ndarray) and g.shape == (cfg["H"],)
    assert abs(c - gold["cost"][1]) < ATOL
    np.testing.assert_allclose(g, gold["grad"][1], rtol=0, atol=ATOL_GRAD)
    np.testing.assert_allclose(ctrl.states_mu_pred.numpy(), gold["states_mu_pred"][1], rtol=0, atol=ATOL)  # :279-283
    np.testing.assert_allclose(ctrl.rewards_traj_var.numpy(), gold["rewards_traj_var"][1], rtol=0, atol=ATOL)
    assert abs(ctrl.cost_traj_mean_lcb.item() - gold["cost_traj_mean_lcb"][1]) < ATOL
    costs, grads = ctrl.compute_mean_lcb_trajectory_batch(cfg["actions"].reshape(3, -1), obs_mu, obs_var)
    np.testing.assert_allclose(costs.cpu().numpy(), gold["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(grads.cpu().numpy(), gold["grad"], rtol=0, atol=ATOL_GRAD)
    # full control step through scipy L-BFGS-B (gp_mpc_controller.py:52-153)
    ctrl.memory.model_inputs[:cfg["N"]] = torch.as_tensor(cfg["x"])
    ctrl.memory.model_targets[:cfg["N"]] = torch.as_tensor(cfg["y"])
    ctrl.memory.len_mem_model = cfg["N"]
    act = ctrl.get_action(np.array([0.1, -0.2, 0.3]))
    assert act.shape == (1,) and -1.0 <= act[0] <= 1.0
    info = ctrl.get_iter_info()
    assert info.predicted_states.shape == (cfg["H"] + 1, 3) and np.isfinite(info.lower_bound_mean_predicted_cost)


@pytest.mark.parametrize("method,iters", [("lbfgs", 15), ("adam", 40)])
def test_batched_on_device_optimizer_beats_serial_restarts(method, iters):
    """B candidates optimised simultaneously (one batched rollout per iteration) reach a cost at
    least as good as the reference-style serial scipy L-BFGS-B restarts from the same warm start."""
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.actions_config import ActionsConfig
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.config_classes.observation_config import ObservationConfig
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.config_classes.total_config import Config
    cfg = make_workload("C2", B=1, H=8, seed=31, N=120)
    r = cfg["reward"]

    def controller(**ctl_kwargs):
        config = Config(
            observation_config=ObservationConfig(obs_var_norm=[cfg["obs_var"]] * 3),
            reward_config=RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                                       weight_state_terminal=list(r["weight_state_terminal"]),
                                       target_action_norm=list(r["target_action"]), weight_action=list(r["weight_action"]),
                                       exploration_factor=r["exploration_factor"]),
            actions_config=ActionsConfig(), controller_config=ControllerConfig(len_horizon=cfg["H"], **ctl_kwargs),
            model_config=ModelConfig(gp_init={"noise_covar.noise": list(cfg["noise"]),
                                              "base_kernel.lengthscale": [list(v) for v in cfg["lengthscale"]],
                                              "outputscale": list(cfg["outputscale"])},
                                     min_std_noise=1e-4, max_std_noise=1.0, min_outputscale=1e-6, max_outputscale=10.0,
                                     min_lengthscale=1e-3, max_lengthscale=1e3))
        c = GpMpcController(-np.ones(3), np.ones(3), -np.ones(1), np.ones(1), config)
        c.memory.model_inputs[:cfg["N"]] = torch.as_tensor(cfg["x"])
        c.memory.model_targets[:cfg["N"]] = torch.as_tensor(cfg["y"])
        c.memory.len_mem_model = cfg["N"]
        return c

    obs = np.array([0.1, -0.2, 0.3])
    np.random.seed(0)
    torch.manual_seed(0)
    serial = controller(restarts_optim=2)
    a_serial = serial.get_action(obs)
    cost_serial = serial.last_optim_cost
    batched = controller(batched_candidates=256, batched_iters=iters, batched_lr=0.05, batched_method=method)
    a_batched = batched.get_action(obs)
    cost_batched = batched.last_optim_cost
    assert abs(cost_batched + batched.cost_traj_mean_lcb.item()) < 1e-8      # side effects describe the winner
    assert a_batched.shape == a_serial.shape and -1.0 <= a_batched[0] <= 1.0
    assert np.isfinite(cost_batched) and cost_batched <= cost_serial + 1e-6
    assert batched.batched_costs.shape == (256,) and batched.get_iter_info().predicted_states.shape == (9, 3)


def test_kernel_path_selection_follows_the_hyperparameters():
    """prepare() re-detects whether all GPs share their hyper-parameters; both paths give the same answer."""
    cfg = make_workload("C3", B=4, H=3, seed=41, N=100)
    eng = make_engine(cfg)
    assert eng.uses_uniform_path()
    uni = rollout(eng, cfg)
    eng.set_path(1)
    assert not eng.uses_uniform_path()
    gen = rollout(eng, cfg)
    np.testing.assert_allclose(uni["cost"], gen["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uni["grad"], gen["grad"], rtol=0, atol=ATOL_GRAD)
    eng.set_path(0)
    ls = full_lengthscale(cfg).copy()
    ls[1, 0] *= 1.0000001                                   # one GP differs in one lengthscale -> general path
    eng.prepare(cfg["x"], cfg["y"], ls, cfg["outputscale"], cfg["noise"])
    assert not eng.uses_uniform_path()
    eng.prepare(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"])
    assert eng.uses_uniform_path()


@pytest.mark.parametrize("kw", [dict(E=4, Na=2, N=200, H=6, B=3, ls=0.3, seed=61), dict(E=8, Na=3, N=130, H=4, B=2, ls=0.7, seed=62)])
def test_reverse_sweep_records_in_global_scratch_match_shared_memory(kw, monkeypatch):
    """The per-step small matrices / stage-cost adjoints of the reverse sweep are precomputed per candidate either into
    shared memory or (shapes whose plan is tight: config 5) into a per-CTA global scratch, or recomputed per step: same
    arithmetic, same gradient."""
    cfg = make_workload(**kw)
    out = {}
    for mode in ("1", "2", "0"):
        monkeypatch.setenv("GPMPC_UNI_PREMAT", mode)
        eng = make_engine(cfg)
        assert eng.uses_uniform_path()
        out[mode] = rollout(eng, cfg)
    np.testing.assert_allclose(out["2"]["cost"], out["1"]["cost"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["2"]["grad"], out["1"]["grad"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["0"]["grad"], out["1"]["grad"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("kw", [dict(E=2, Na=1, N=150, H=4, B=5, ls=0.5, seed=71), dict(E=3, Na=1, N=200, H=3, B=4, ls=0.5, seed=72),
                                dict(E=4, Na=2, N=260, H=3, B=3, ls=0.3, seed=73),
                                dict(E=5, Na=2, N=130, H=3, B=2, ls=0.6, seed=74, include_time_model=True)])
def test_both_builds_of_the_uniform_kernels_agree(kw, monkeypatch):
    """The uniform kernels exist in two builds for E <= 5 -- 256 threads at <= 128 registers and 128 threads at <= 168
    (three CTAs per SM; the host takes it for large batches) -- and must return the same costs and gradients."""
    cfg = make_workload(**kw)
    out = {}
    for thr in ("256", "128"):
        monkeypatch.setenv("GPMPC_UNI_FWD_THREADS", thr)
        monkeypatch.setenv("GPMPC_UNI_BWD_THREADS", thr)
        monkeypatch.setenv("GPMPC_UNI_CLUSTER", "1")
        eng = make_engine(cfg)
        assert eng.uses_uniform_path()
        out[thr] = rollout(eng, cfg)
    # (4 instead of 8 warps split the sums differently; the 1e6..1e8 cancellation of the covariance sums turns that
    # into ~1e-10, as between any two schedules of the same build)
    np.testing.assert_allclose(out["128"]["cost"], out["256"]["cost"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out["128"]["grad"], out["256"]["grad"], rtol=0, atol=1e-8)
    want = orc.evaluate_workload(cfg)
    np.testing.assert_allclose(out["128"]["grad"], want["grad"], rtol=0, atol=ATOL_GRAD)


def test_uniform_path_large_state_dimension():
    """C5 dims (E=8, Na=3, D=11) on the uniform path (255-register kernels, owner-lane reductions)."""
    cfg = make_workload(E=8, Na=3, N=200, H=2, B=2, ls=0.7, seed=42)
    eng = make_engine(cfg)
    assert eng.uses_uniform_path()
    want = orc.evaluate_workload(cfg)
    got = rollout(eng, cfg)
    np.testing.assert_allclose(got["cost"], want["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["grad"], want["grad"], rtol=0, atol=ATOL_GRAD)
    np.testing.assert_allclose(got["states_var_pred"], want["states_var_pred"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("kw", [dict(E=4, Na=2, N=500, H=6, B=1, ls=0.25), dict(E=3, Na=1, N=300, H=5, B=5, ls=0.5),
                                dict(E=2, Na=1, N=200, H=7, B=20, ls=0.5, limit_action_change=True),
                                dict(E=5, Na=2, N=260, H=4, B=2, ls=0.6, include_time_model=True)])
def test_cluster_mode_matches_single_cta_mode(kw, monkeypatch):
    """Small batches run with a thread-block cluster per candidate (sweep split over 2/4/8 CTAs, sums met in L2, one
    cluster barrier per step).  Same numbers as the one-CTA-per-candidate path and as the oracle."""
    cfg = make_workload(seed=11, **kw)
    eng = make_engine(cfg)
    assert eng.uses_uniform_path()
    got = rollout(eng, cfg)                      # automatic: clusters (B * cluster <= number of SMs)
    got2 = rollout(eng, cfg)                     # again: the rotating accumulators must come back clean
    monkeypatch.setenv("GPMPC_UNI_CLUSTER", "1")
    plain = rollout(eng, cfg)
    monkeypatch.setenv("GPMPC_UNI_CLUSTER", "2")
    two = rollout(eng, cfg)
    monkeypatch.delenv("GPMPC_UNI_CLUSTER")
    for other in (got2, plain, two):   # not bit-for-bit: the float64 partial sums meet in scheduling order
        for key in ("cost", "grad", "states_mu_pred", "states_var_pred", "rewards_trajectory", "rewards_traj_var"):
            np.testing.assert_allclose(got[key], other[key], rtol=0, atol=ATOL_GRAD if key == "grad" else ATOL, err_msg=key)
    want = orc.evaluate_workload(cfg, candidates=[0])
    np.testing.assert_allclose(got["cost"][:1], want["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["grad"][:1], want["grad"].reshape(1, -1), rtol=0, atol=ATOL_GRAD)
    fwd_only = rollout(eng, cfg, need_grad=False)
    np.testing.assert_allclose(fwd_only["cost"], got["cost"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("kw", [dict(E=4, Na=2, N=300, H=5, B=1, ls=0.25), dict(E=3, Na=1, N=200, H=4, B=7, ls=0.5),
                                dict(E=2, Na=1, N=130, H=6, B=30, ls=0.5, limit_action_change=True, use_constraints=True),
                                dict(E=5, Na=2, N=140, H=3, B=2, ls=0.6, include_time_model=True)])
def test_general_path_cluster_mode_matches_single_cta_mode(kw, monkeypatch):
    """Per-GP hyper-parameters, small batch: the output pairs (a, b) are dealt to the CTAs of a thread-block cluster and
    their S_raw exchanged through L2 around one cluster barrier per step.  Same numbers as the one-CTA path / oracle."""
    cfg = make_workload(seed=12, distinct_lengthscales=True, **kw)
    eng = make_engine(cfg)
    assert not eng.uses_uniform_path()
    got = rollout(eng, cfg)
    got2 = rollout(eng, cfg)
    monkeypatch.setenv("GPMPC_GEN_CLUSTER", "1")
    plain = rollout(eng, cfg)
    monkeypatch.setenv("GPMPC_GEN_CLUSTER", "2")
    two = rollout(eng, cfg)
    monkeypatch.delenv("GPMPC_GEN_CLUSTER")
    for other in (got2, plain, two):   # not bit-for-bit: the float64 partial sums meet in scheduling order
        for key in ("cost", "grad", "states_mu_pred", "states_var_pred", "rewards_trajectory", "rewards_traj_var"):
            np.testing.assert_allclose(got[key], other[key], rtol=0, atol=ATOL_GRAD if key == "grad" else ATOL, err_msg=key)
    want = orc.evaluate_workload(cfg, candidates=[0])
    np.testing.assert_allclose(got["cost"][:1], want["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(got["grad"][:1], want["grad"].reshape(1, -1), rtol=0, atol=ATOL_GRAD)
    fwd_only = rollout(eng, cfg, need_grad=False)
    np.testing.assert_allclose(fwd_only["cost"], got["cost"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("n,history,iters", [(60, 8, 25), (7, 3, 12), (200, 5, 10)])
def test_fused_lbfgs_update_kernel_follows_its_torch_specification(n, history, iters):
    """gpmpc_lbfgs_update (csrc/gpmpc_optim.cu: one kernel per optimiser iteration) against
    batched_optim.minimize_box_lbfgs (plain torch) on a batch of box-constrained problems with active bounds, a
    non-quadratic term, a candidate whose objective turns non-finite off the start point and NaN gradient entries:
    same iterates."""
    from rl_gp_mpc.control_objects.controllers.batched_optim import minimize_box_lbfgs
    g = torch.Generator().manual_seed(5)
    nb = 37
    dev = torch.device("cuda")
    tgt = (torch.rand((nb, n), generator=g, dtype=torch.float64) * 1.6 - 0.3).to(dev)     # some optima outside the box
    w = (0.5 + 4.0 * torch.rand((nb, n), generator=g, dtype=torch.float64)).to(dev)
    x0 = torch.rand((nb, n), generator=g, dtype=torch.float64).to(dev)
    x0[3, :] = 0.0                                                                          # starts on a bound

    def fun(x):
        d = x - tgt
        cost = (w * d * d).sum(1) + 0.1 * torch.sin(3.0 * x).sum(1)
        grad = 2.0 * w * d + 0.3 * torch.cos(3.0 * x)
        bad = (x[:, 0] > 0.9)                                                               # non-finite region for some candidates
        cost = torch.where(bad & (torch.arange(nb, device=dev) % 5 == 1), torch.full_like(cost, float("nan")), cost)
        grad = grad.clone()
        grad[7, 1] = float("nan")
        return cost, grad

    xa, fa = minimize_box_lbfgs(fun, x0, iters, history=history)
    xb, fb = minimize_box_lbfgs(fun, x0, iters, history=history, fused=True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(xb.cpu().numpy(), xa.cpu().numpy(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(fb.cpu().numpy(), fa.cpu().numpy(), rtol=0, atol=1e-6)
    assert float((fa - fun(x0)[0]).nan_to_num(0.0).max()) <= 0.0                            # never worse than the start


def test_c5_full_size_gradient_general_path_agrees_with_uniform_path():
    """BASELINE.json config 5 dims at the full training-set size (E=8, D=11, N=1000) WITH the gradient on the general
    (per-GP hyper-parameter) kernel: its shared-memory plan only fits with the per-GP weights of the mean part in the
    global scratch (gpmpc_api.cu, lb_global).  With identical hyper-parameters the general and the uniform kernels
    compute the same function by different algorithms (forward-mode Jacobians vs reverse sweep): they must agree."""
    cfg = make_workload("C5", B=2, H=2, seed=55)
    uni = rollout(make_engine(cfg, 0), cfg)
    gen = rollout(make_engine(cfg, 1), cfg)
    np.testing.assert_allclose(gen["cost"], uni["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(gen["grad"], uni["grad"], rtol=0, atol=ATOL_GRAD)
    np.testing.assert_allclose(gen["states_var_pred"], uni["states_var_pred"], rtol=0, atol=ATOL)
    assert np.abs(uni["grad"]).max() > 1e-6
