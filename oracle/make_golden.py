"""Generate tests/golden/*.npz from the reference's OWN code -- run in the build container only.

    python -m oracle.make_golden

Imports /root/reference verbatim through oracle/ref_loader.py and records, for a
set of small seeded workloads that exercise every toggle of the hot path
(SURVEY.md section 4), the outputs of
  calculate_factorizations            gp_model.py:400
  predict_next_state_change           gp_model.py:112
  predict_trajectory                  gp_model.py:60
  get_rewards_trajectory              setpoint_distance_reward_mapper.py:144
  compute_mean_lcb_trajectory         gp_mpc_controller.py:229   (value + autograd gradient)
The vectors are the pin for oracle/gpmpc_oracle.py (CPU tests) and for the CUDA
path (GPU tests); /root/reference is never read at test time on the GPU box.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference, make_reference_controller  # noqa: E402
from oracle.workloads import make_workload  # noqa: E402

CASES = {
    # name: make_workload kwargs
    "pendulum_c1": dict(name="C1", B=3, seed=1),
    "time_model": dict(E=2, Na=2, N=40, H=6, B=2, ls=0.25, preset="process", seed=2,
                       include_time_model=True, iter_ctrl=57),
    "derivative_actions": dict(E=2, Na=1, N=30, H=5, B=3, ls=0.5, preset="mountaincar", seed=3,
                               limit_action_change=True),
    "constraints_clip": dict(E=3, Na=1, N=30, H=5, B=2, ls=0.5, preset="pendulum", seed=4,
                             use_constraints=True, clip_lower_bound_cost_to_0=True),
    "distinct_ls_e4": dict(E=4, Na=2, N=60, H=4, B=2, ls=0.4, seed=5, distinct_lengthscales=True),
    "single_state": dict(E=1, Na=1, N=25, H=4, B=2, ls=0.5, seed=6),
    "large_obs_var": dict(E=3, Na=2, N=45, H=5, B=2, ls=0.5, seed=7, obs_var=3e-3, noise=1e-3,
                          exploration_factor=3.0),
    "empty_memory": dict(E=2, Na=1, N=1, H=3, B=2, ls=0.5, seed=8),
}

# The reference's own hyper-parameter regime at the BASELINE.json training-set sizes: noise 1e-5 at N = 200 / 500 gives
# cond(K + noise I) ~ 1e6 .. 1e7, where the covariance sums cancel by ~1e8.  `python -m
# oracle.make_golden --large` writes tests/golden/big_<name>.npz WITHOUT the (E, N, N) inverse (4 MB at N = 500):
# its diagonal and row sums are kept instead, beta and every output of the path in full.
LARGE_CASES = {
    "c4a_n500": dict(name="C4a", B=1, H=3, seed=41),                       # E=2, Na=2, N=500 (ProcessControl dims)
    "c2_n200_h25": dict(name="C2", B=1, H=25, seed=42),                    # E=3, Na=1, N=200, full horizon
    "c4a_n500_distinct": dict(name="C4a", B=1, H=3, seed=43, distinct_lengthscales=True),   # general kernel path
}


def run_case(kwargs, ref, large=False):
    cfg = make_workload(**kwargs)
    if kwargs.get("name") is None and cfg["N"] == 1:
        cfg["x"][:] = 0.0  # gp_memory.py:109-111: zeros (1,D)/(1,E) when the memory is empty
        cfg["y"][:] = 0.0
    ctrl = make_reference_controller(cfg, ref)
    tm = ctrl.transition_model
    x = torch.as_tensor(cfg["x"]); y = torch.as_tensor(cfg["y"])
    with torch.no_grad():
        tm.prepare_inference(x, y)
    out = dict(iK=tm.iK.numpy(), beta=tm.beta.numpy(),
               lengthscales=tm.lengthscales.detach().numpy(), variances=tm.variances.detach().numpy())
    if large:
        iK = out.pop("iK")
        out.update(iK_diag=np.diagonal(iK, axis1=1, axis2=2).copy(), iK_rowsum=iK.sum(axis=2),
                   iK_absmax=np.abs(iK).max())
    # one moment-matching step at a full (non-diagonal) input covariance
    E, D = cfg["E"], cfg["D"]
    rng = np.random.default_rng(77 + cfg["seed"])
    A = rng.standard_normal((E, E)) * 0.03
    s = A @ A.T + 1e-4 * np.eye(E)
    in_var = np.zeros((D, D)); in_var[:E, :E] = s
    in_mu = rng.uniform(0.2, 0.8, size=(D,))
    if cfg["include_time_model"]:
        in_mu[-1] = cfg["iter_ctrl"] + 2
    with torch.no_grad():
        M, S, V = tm.predict_next_state_change(torch.as_tensor(in_mu), torch.as_tensor(in_var))
    out.update(step_in_mu=in_mu, step_in_var=in_var, step_M=M.numpy(), step_S=S.numpy(), step_V=V.numpy())
    obs_mu = torch.as_tensor(cfg["mu0"]); obs_var = torch.as_tensor(cfg["Sigma0"])
    costs, grads, mus, vars_, rew, rewv, lcb = [], [], [], [], [], [], []
    for b in range(cfg["B"]):
        c, g = ctrl.compute_mean_lcb_trajectory(cfg["actions"][b].reshape(-1).copy(), obs_mu, obs_var)
        costs.append(c); grads.append(g)
        mus.append(ctrl.states_mu_pred.numpy()); vars_.append(ctrl.states_var_pred.numpy())
        rew.append(ctrl.rewards_trajectory.numpy()); rewv.append(ctrl.rewards_traj_var.numpy())
        lcb.append(ctrl.cost_traj_mean_lcb.item())
    out.update(cost=np.array(costs), grad=np.stack(grads), states_mu_pred=np.stack(mus),
               states_var_pred=np.stack(vars_), rewards_trajectory=np.stack(rew),
               rewards_traj_var=np.stack(rewv), cost_traj_mean_lcb=np.array(lcb))
    out["x"] = cfg["x"]; out["y"] = cfg["y"]; out["actions"] = cfg["actions"]; out["mu0"] = cfg["mu0"]
    return out


def main():
    ref = load_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    if "--large" in sys.argv:
        with open(os.path.join(outdir, "cases_big.json"), "w") as f:
            json.dump(LARGE_CASES, f, indent=1, sort_keys=True)
        for name, kwargs in LARGE_CASES.items():
            out = run_case(kwargs, ref, large=True)
            np.savez_compressed(os.path.join(outdir, "big_" + name + ".npz"), **out)
            print(name, "cost", out["cost"], "|grad|max", np.abs(out["grad"]).max(), "|iK|max", out["iK_absmax"])
        return
    with open(os.path.join(outdir, "cases.json"), "w") as f:
        json.dump(CASES, f, indent=1, sort_keys=True)
    for name, kwargs in CASES.items():
        out = run_case(kwargs, ref)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print(name, "cost", out["cost"], "|grad|max", np.abs(out["grad"]).max())


if __name__ == "__main__":
    main()
