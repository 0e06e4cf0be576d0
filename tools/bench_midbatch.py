"""Latency of ONE batched objective+gradient evaluation for small and mid-size candidate batches (the batched optimiser's
regime): thread-block clusters share a candidate while clusters x CTAs fit the device.
    python tools/bench_midbatch.py [workload] [--distinct]          (GPMPC_UNI_CLUSTER_CAP=1: one CTA per SM, round-1 rule)"""
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


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "C4b"
    distinct = "--distinct" in sys.argv
    cfg = make_workload(name, B=512, distinct_lengthscales=distinct)
    eng = _cabi.Engine("cuda:0")
    eng.prepare(cfg["x"], cfg["y"], full_lengthscale(cfg), cfg["outputscale"], cfg["noise"])
    r = cfg["reward"]
    eng.set_cost(np.concatenate([r["target_state"], r["target_action"]]).astype(float),
                 np.diag(np.concatenate([r["weight_state"], r["weight_action"]]).astype(float)),
                 np.diag(np.asarray(r["weight_state_terminal"], float)), r["exploration_factor"])
    eng.enable_timing(True)
    a_all = torch.as_tensor(cfg["actions"].reshape(512, -1)).cuda()
    mu, s0 = torch.as_tensor(cfg["mu0"]).cuda(), torch.as_tensor(cfg["Sigma0"]).cuda()
    print("# %s%s N=%d H=%d, one objective+gradient evaluation; cap rule: %s" % (
        name, " distinct hyper-parameters" if distinct else "", cfg["N"], cfg["H"], os.environ.get("GPMPC_UNI_CLUSTER_CAP", "2 (default)")))
    for B in (1, 8, 16, 32, 64, 96, 128, 148, 192, 256, 296, 512):
        a = a_all[:B].contiguous()
        for _ in range(3):
            eng.rollout(a, mu, s0, cfg["H"], need_grad=True, need_traj=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            eng.rollout(a, mu, s0, cfg["H"], need_grad=True, need_traj=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps * 1e3
        print("B=%4d  %.3f ms per evaluation  (kernels: fwd %.3f + reverse %.3f ms)  %.0f predictions/s" % (
            B, dt, eng.last_rollout_ms(), max(eng.last_backward_ms(), 0.0), B * cfg["H"] / dt * 1e3))


if __name__ == "__main__":
    main()
