"""Measurements for the SURVEY.md 8(f) rows around the hot path (one B200):
  N1  one control step (GpMpcController.get_action): serial scipy L-BFGS-B restarts vs the batched on-device optimiser
  N2  hyper-parameter fit (GpStateTransitionModel.train) on the device objective: wall time, evaluations per second
  N3  growing the factorisation by one point: gpmpc_append vs gpmpc_prepare
Prints one line per measurement."""
import os
import queue
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200"))

from oracle.workloads import make_workload  # noqa: E402
from rl_gp_mpc import GpMpcController, _cabi  # noqa: E402
from rl_gp_mpc.config_classes.actions_config import ActionsConfig  # noqa: E402
from rl_gp_mpc.config_classes.controller_config import ControllerConfig  # noqa: E402
from rl_gp_mpc.config_classes.model_config import ModelConfig  # noqa: E402
from rl_gp_mpc.config_classes.observation_config import ObservationConfig  # noqa: E402
from rl_gp_mpc.config_classes.reward_config import RewardConfig  # noqa: E402
from rl_gp_mpc.config_classes.total_config import Config  # noqa: E402
from rl_gp_mpc.config_classes.training_config import TrainingConfig  # noqa: E402
from rl_gp_mpc.control_objects.models.gp_model import GpStateTransitionModel  # noqa: E402


def controller(cfg, batched=0, restarts=2, method="lbfgs", iters=30):
    E, Na, H = cfg["E"], cfg["Na"], cfg["H"]
    r = cfg["reward"]
    config = Config(
        observation_config=ObservationConfig(obs_var_norm=[cfg["obs_var"]] * E),
        reward_config=RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                                   weight_state_terminal=list(r["weight_state_terminal"]),
                                   target_action_norm=list(r["target_action"]), weight_action=list(r["weight_action"]),
                                   exploration_factor=r["exploration_factor"]),
        actions_config=ActionsConfig(),
        controller_config=ControllerConfig(len_horizon=H, restarts_optim=restarts, batched_candidates=batched,
                                           batched_iters=iters, batched_method=method),
        training_config=TrainingConfig(training_frequency=10 ** 9),
        model_config=ModelConfig(gp_init={"noise_covar.noise": list(cfg["noise"]),
                                          "base_kernel.lengthscale": [list(v) for v in cfg["lengthscale"]],
                                          "outputscale": list(cfg["outputscale"])},
                                 min_std_noise=1e-4, max_std_noise=1.0, min_outputscale=1e-6, max_outputscale=10.0,
                                 min_lengthscale=1e-3, max_lengthscale=1e3))
    ctrl = GpMpcController(np.zeros(E), np.ones(E), np.zeros(Na), np.ones(Na), config)
    m = ctrl.memory                      # fill the replay store with the workload's transitions
    n = len(cfg["x"])
    m.model_inputs = torch.as_tensor(cfg["x"]).clone()
    m.model_targets = torch.as_tensor(cfg["y"]).clone()
    m.len_mem_model = n
    return ctrl


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    for name in ("C2", "C4b"):
        cfg = make_workload(name, B=1)
        obs = cfg["mu0"]
        for label, kw in (("scipy L-BFGS-B, 2 restarts", dict(batched=0, restarts=2)),
                          ("batched on device, Adam, 256 candidates x 30 iterations", dict(batched=256, method="adam")),
                          ("batched on device, L-BFGS, 256 candidates x 15 iterations", dict(batched=256, iters=15)),
                          ("batched on device, L-BFGS, 64 candidates x 15 iterations", dict(batched=64, iters=15)),
                          ("batched on device, L-BFGS, 16 candidates x 15 iterations", dict(batched=16, iters=15))):
            ctrl = controller(cfg, **kw)
            ctrl.get_action(obs)         # warm-up (first call: library load, allocations)
            t0 = time.perf_counter()
            reps = 3
            costs = []
            for _ in range(reps):
                ctrl.actions_mpc_previous_iter = None
                ctrl.get_action(obs)
                costs.append(ctrl.last_optim_cost)
            dt = (time.perf_counter() - t0) / reps
            print("N1 %-4s N=%d H=%d  %-62s %8.1f ms per control step, objective reached %.6f" % (
                name, cfg["N"], cfg["H"], label, dt * 1e3, float(np.mean(costs))), flush=True)
    # N2: hyper-parameter fit on the device objective
    for name in ("C2", "C4b"):
        cfg = make_workload(name, B=1)
        ctrl = controller(cfg)
        tm = ctrl.transition_model
        tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
        st = tm.save_state()
        st.to_arrays()
        q = queue.Queue()
        eng = tm.engine
        l0 = eng.launch_count()
        for lockstep in (False, True):     # same random restarts for both procedures
            for rep in range(2):
                torch.manual_seed(1234)
                t0 = time.perf_counter()
                GpStateTransitionModel.train(q, st, 7e-3, 15, 1e-3, False, 5, lockstep=lockstep)
                dt = time.perf_counter() - t0
                q.get()
            print("N2 %-4s N=%d E=%d  train(): 15 LBFGS iterations per GP, all GPs, %-38s %.2f s (second call)" % (
                name, cfg["N"], cfg["E"], "lockstep-batched evaluations:" if lockstep else "one GP after the other:", dt), flush=True)
    # N3: append vs prepare
    for name in ("C2", "C4b", "C5"):
        cfg = make_workload(name, B=1, H=2)
        eng = _cabi.Engine()
        from oracle.workloads import full_lengthscale
        ls = full_lengthscale(cfg)
        x = torch.as_tensor(cfg["x"]).cuda(); y = torch.as_tensor(cfg["y"]).cuda()
        n = cfg["N"] - 1
        eng.prepare(x, y, ls, cfg["outputscale"], cfg["noise"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            eng.prepare(x, y, ls, cfg["outputscale"], cfg["noise"])
        torch.cuda.synchronize()
        tp = (time.perf_counter() - t0) / 5
        ta = []
        for _ in range(5):
            eng.prepare(x[:n], y[:n], ls, cfg["outputscale"], cfg["noise"])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.append(x[n], y[n])
            torch.cuda.synchronize()
            ta.append(time.perf_counter() - t0)
        print("N3 %-4s N=%d E=%d  gpmpc_prepare %.3f ms   gpmpc_append %.3f ms" % (
            name, cfg["N"], cfg["E"], tp * 1e3, float(np.median(ta)) * 1e3), flush=True)


if __name__ == "__main__":
    main()
