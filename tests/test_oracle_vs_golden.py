"""CPU: pin oracle/gpmpc_oracle.py against vectors produced by the reference's own code."""
import numpy as np
import pytest
import torch

from oracle import gpmpc_oracle as orc
from tests.golden_utils import case_names, load_case


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference_golden(name):
    cfg, gold = load_case(name)
    model = orc.model_from_workload(cfg)
    # factorisation (gp_model.py:400): iK entries reach 1e5, compare relative to scale
    scale = np.abs(gold["iK"]).max()
    assert np.abs(model.iK.numpy() - gold["iK"]).max() <= 1e-7 * scale
    assert np.abs(model.beta.numpy() - gold["beta"]).max() <= 1e-7 * max(1.0, np.abs(gold["beta"]).max())
    # one moment-matching step (gp_model.py:112)
    M, S, V = orc.predict_next_state_change(model, torch.as_tensor(gold["step_in_mu"]),
                                            torch.as_tensor(gold["step_in_var"]))
    np.testing.assert_allclose(M.numpy(), gold["step_M"], rtol=0, atol=2e-9)
    np.testing.assert_allclose(S.numpy(), gold["step_S"], rtol=0, atol=2e-9)
    np.testing.assert_allclose(V.numpy(), gold["step_V"], rtol=0, atol=2e-8)
    # objective, gradient and the five side-effect tensors (gp_mpc_controller.py:229-285)
    res = orc.evaluate_workload(cfg, model=model)
    np.testing.assert_allclose(res["cost"], gold["cost"], rtol=0, atol=2e-9)
    np.testing.assert_allclose(res["grad"], gold["grad"], rtol=0, atol=2e-8)
    np.testing.assert_allclose(res["states_mu_pred"], gold["states_mu_pred"], rtol=0, atol=2e-9)
    np.testing.assert_allclose(res["states_var_pred"], gold["states_var_pred"], rtol=0, atol=2e-9)
    np.testing.assert_allclose(res["rewards_trajectory"], gold["rewards_trajectory"], rtol=0, atol=2e-9)
    np.testing.assert_allclose(res["rewards_traj_var"], gold["rewards_traj_var"], rtol=0, atol=2e-9)
    np.testing.assert_allclose(res["cost_traj_mean_lcb"], gold["cost_traj_mean_lcb"], rtol=0, atol=2e-9)
