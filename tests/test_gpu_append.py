"""gpmpc_append: growing the factorisation one point at a time must give what a fresh
gpmpc_prepare on the grown training set gives -- same iK / beta (to the rounding the condition number allows) and the
same rollout costs and gradients within the stated parity tolerance -- on both kernel paths; and it must refuse
(without touching the factorisation) once the padded size is used up."""
import numpy as np
import pytest
import torch

from oracle import gpmpc_oracle as orc
from oracle.workloads import full_lengthscale, make_workload
from tests.test_gpu_parity import ATOL, ATOL_GRAD, make_engine, rollout

pytestmark = pytest.mark.gpu


def _grown(cfg, n0):
    part = dict(cfg)
    part["x"] = cfg["x"][:n0].copy()
    part["y"] = cfg["y"][:n0].copy()
    part["N"] = n0
    return part


@pytest.mark.parametrize("distinct", [False, True])
@pytest.mark.parametrize("n0,n1", [(130, 150), (37, 64), (1, 9)])
def test_append_matches_fresh_prepare(n0, n1, distinct):
    cfg = make_workload(E=3, Na=2, N=n1, H=8, B=6, ls=0.4, seed=5, distinct_lengthscales=distinct)
    eng = make_engine(_grown(cfg, n0))
    assert eng.append_room() == (n0 + 63) // 64 * 64 - n0
    for i in range(n0, n1):
        eng.append(cfg["x"][i], cfg["y"][i])
    assert eng.N == n1
    fresh = make_engine(cfg)
    iK_a, beta_a = (t.cpu().numpy() for t in eng.factorization())
    iK_f, beta_f = (t.cpu().numpy() for t in fresh.factorization())
    scale = np.abs(iK_f).max()
    assert np.abs(iK_a - iK_f).max() <= 1e-8 * scale
    assert np.abs(beta_a - beta_f).max() <= 1e-8 * max(1.0, np.abs(beta_f).max())
    out_a, out_f = rollout(eng, cfg), rollout(fresh, cfg)
    np.testing.assert_allclose(out_a["cost"], out_f["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(out_a["grad"], out_f["grad"], rtol=0, atol=ATOL_GRAD)
    np.testing.assert_allclose(out_a["states_var_pred"], out_f["states_var_pred"], rtol=0, atol=ATOL)
    # and against the CPU oracle on the grown set
    ref = orc.evaluate_workload(cfg)
    np.testing.assert_allclose(out_a["cost"], ref["cost"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(out_a["grad"], ref["grad"].reshape(out_a["grad"].shape), rtol=0, atol=ATOL_GRAD)


def test_append_keeps_the_marginal_likelihood_consistent():
    cfg = make_workload(E=2, Na=1, N=90, H=4, B=2, ls=0.5, seed=2)
    eng = make_engine(_grown(cfg, 70))
    for i in range(70, 90):
        eng.append(cfg["x"][i], cfg["y"][i])
    fresh = make_engine(cfg)
    np.testing.assert_allclose(eng.mll(cfg["y"]).cpu().numpy(), fresh.mll(cfg["y"]).cpu().numpy(), rtol=1e-8, atol=1e-7)


def test_append_refuses_when_the_padded_size_is_used_up():
    from rl_gp_mpc import _cabi
    cfg = make_workload(E=2, Na=1, N=64, H=4, B=2, ls=0.5, seed=3)
    eng = make_engine(cfg)
    assert eng.append_room() == 0
    before = eng.factorization()[0].clone()
    with pytest.raises(_cabi.GpmpcError, match="padded size"):
        eng.append(cfg["x"][0] + 0.01, cfg["y"][0])
    assert eng.N == 64
    assert torch.equal(before, eng.factorization()[0])


def test_model_prepare_inference_appends_when_the_memory_grew():
    """GpStateTransitionModel.prepare_inference(inputs, state_changes) takes the O(N^2) path when called with the
    previous training set plus new rows (what GpMpcController does after Memory.add), else refactorises."""
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.control_objects.models.gp_model import GpStateTransitionModel
    cfg = make_workload(E=2, Na=1, N=80, H=5, B=3, ls=0.5, seed=4)
    x, y = torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"])
    def model_config():
        return ModelConfig(gp_init={"noise_covar.noise": [1e-4] * 2, "base_kernel.lengthscale": [[0.75] * 3] * 2,
                                    "outputscale": [5e-2] * 2})
    model = GpStateTransitionModel(model_config(), dim_state=2, dim_action=1)
    model.prepare_inference(x[:70], y[:70])
    assert model.last_prepare_mode == "full"
    model.prepare_inference(x[:73], y[:73])
    assert model.last_prepare_mode == "append" and model.engine.N == 73
    iK_a, beta_a = model.engine.factorization()
    ref = GpStateTransitionModel(model_config(), dim_state=2, dim_action=1)
    ref.incremental_updates = False
    ref.prepare_inference(x[:73], y[:73])
    assert ref.last_prepare_mode == "full"
    iK_f, beta_f = ref.engine.factorization()
    assert float((iK_a - iK_f).abs().max()) <= 1e-8 * float(iK_f.abs().max())
    model.prepare_inference(x[:80], y[:80])                   # 7 more rows, still inside the padded size of 128
    assert model.last_prepare_mode == "append"
    y2 = y.clone(); y2[0, 0] += 1.0
    model.prepare_inference(x[:80], y2[:80])                  # same size, different data: must refactorise
    assert model.last_prepare_mode == "full"
