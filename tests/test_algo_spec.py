"""CPU: the factorised algorithm + hand-derived adjoint (tests/algo_spec.py, the blueprint of the
CUDA kernels) reproduces the reference golden vectors."""
import numpy as np
import pytest

from oracle.workloads import full_lengthscale
from tests import algo_spec as sp
from tests.golden_utils import case_names, load_case


@pytest.mark.parametrize("name", case_names())
def test_spec_matches_reference_golden(name):
    cfg, gold = load_case(name)
    d = sp.Data(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"],
                iK=gold["iK"], beta=gold["beta"])
    E, D = cfg["E"], cfg["D"]
    st = sp.step_forward(d, gold["step_in_mu"], gold["step_in_var"][:E, :E])
    np.testing.assert_allclose(st["M"], gold["step_M"][0], rtol=0, atol=5e-9)
    np.testing.assert_allclose(st["S"], gold["step_S"], rtol=0, atol=5e-9)
    np.testing.assert_allclose(st["V"].T, gold["step_V"], rtol=0, atol=5e-8)
    cost = sp.Cost(cfg["reward"], E, cfg["Na"])
    for b in range(cfg["B"]):
        r = sp.rollout(d, cost, cfg["actions"][b], cfg["mu0"], cfg["Sigma0"], cfg["H"], cfg["iter_ctrl"],
                       cfg["include_time_model"], cfg["limit_action_change"], cfg["max_change_action_norm"],
                       cfg["action_prev"])
        np.testing.assert_allclose(r["cost"], gold["cost"][b], rtol=0, atol=5e-9)
        np.testing.assert_allclose(r["states_mu_pred"], gold["states_mu_pred"][b], rtol=0, atol=5e-9)
        np.testing.assert_allclose(r["states_var_pred"], gold["states_var_pred"][b], rtol=0, atol=5e-9)
        np.testing.assert_allclose(r["rewards_traj_var"], gold["rewards_traj_var"][b], rtol=0, atol=5e-9)
        np.testing.assert_allclose(r["grad"], gold["grad"][b], rtol=0, atol=1e-7)
