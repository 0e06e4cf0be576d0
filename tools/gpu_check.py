"""First-light check on the GPU box: errors of every output vs the golden vectors + a quick timing."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200"))

from oracle.workloads import full_lengthscale, make_workload  # noqa: E402
from rl_gp_mpc import _cabi  # noqa: E402
from tests.golden_utils import case_names, load_case  # noqa: E402


def engine_for(cfg, path=0):
    eng = _cabi.Engine()
    eng.set_path(path)
    eng.prepare(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"])
    r = cfg["reward"]
    E, Na = cfg["E"], cfg["Na"]
    W = np.diag(np.concatenate([r["weight_state"], r["weight_action"]]).astype(float))
    WT = np.diag(np.asarray(r["weight_state_terminal"], float))
    tgt = np.concatenate([r["target_state"], r["target_action"]]).astype(float)
    eng.set_cost(tgt, W, WT, r["exploration_factor"], r["use_constraints"], r["state_min"], r["state_max"],
                 r["clip_lower_bound_cost_to_0"])
    return eng


def run_rollout(eng, cfg, need_grad=True, actions=None):
    a = cfg["actions"] if actions is None else actions
    return eng.rollout(a, cfg["mu0"], cfg["Sigma0"], cfg["H"], cfg["iter_ctrl"], cfg["limit_action_change"],
                       cfg["max_change_action_norm"], cfg["action_prev"], need_grad=need_grad)


def err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())


def main():
    print(torch.cuda.get_device_name(0))
    for name in case_names():
      for path in (0, 1):
        cfg, gold = load_case(name)
        eng = engine_for(cfg, path)
        label = name + ("/U" if eng.uses_uniform_path() else "/G")
        iK, beta = eng.factorization()
        E = cfg["E"]
        M, S, V = eng.predict_step(gold["step_in_mu"][None], gold["step_in_var"][None, :E, :E])
        out = run_rollout(eng, cfg)
        out2 = run_rollout(eng, cfg, need_grad=False)
        torch.cuda.synchronize()
        print("%-20s iK %.1e (scale %.1e) beta %.1e | step M %.1e S %.1e V %.1e | cost %.1e grad %.1e mu %.1e var %.1e rv %.1e | fwd-only cost %.1e" % (
            label, err(iK.cpu(), gold["iK"]), np.abs(gold["iK"]).max(), err(beta.cpu(), gold["beta"]),
            err(M.cpu()[0], gold["step_M"][0]), err(S.cpu()[0], gold["step_S"]), err(V.cpu()[0], gold["step_V"]),
            err(out["cost"].cpu(), gold["cost"]), err(out["grad"].cpu(), gold["grad"]),
            err(out["states_mu_pred"].cpu(), gold["states_mu_pred"]), err(out["states_var_pred"].cpu(), gold["states_var_pred"]),
            err(out["rewards_traj_var"].cpu(), gold["rewards_traj_var"]), err(out2["cost"].cpu(), gold["cost"])), flush=True)
    # quick timing on the headline shape (small batch, short horizon)
    for wl, B, H, path in (("C4b", 296, 4, 0), ("C4b", 296, 4, 1), ("C4b", 1184, 6, 0), ("C4a", 296, 4, 0), ("C2", 296, 4, 0), ("C5", 148, 2, 0), ("C5", 148, 2, 1)):
        cfg = make_workload(wl, B=B, H=H)
        eng = engine_for(cfg, path)
        wl = wl + ("/U" if eng.uses_uniform_path() else "/G")
        eng.enable_timing(True)
        for need_grad in (False, True):
            try:
                for it in range(2):
                    run_rollout(eng, cfg, need_grad=need_grad)
                    torch.cuda.synchronize()
            except Exception as exc:
                print("%s B=%d H=%d grad=%d: FAILED %s" % (wl, B, H, need_grad, exc), flush=True)
                continue
            ms, mb = eng.last_rollout_ms(), max(eng.last_backward_ms(), 0.0)
            print("%s B=%d H=%d grad=%d: fwd %.2f ms bwd %.3f ms -> %.0f preds/s" % (
                wl, B, H, need_grad, ms, mb, B * H / (ms + mb) * 1e3), flush=True)
        t0 = time.time(); eng.prepare(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"]); torch.cuda.synchronize()
        print("prepare N=%d: %.1f ms" % (cfg["N"], (time.time() - t0) * 1e3))


if __name__ == "__main__":
    main()
