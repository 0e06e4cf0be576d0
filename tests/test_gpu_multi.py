"""Multi-GPU parity on hardware (needs >= 2 CUDA devices; skipped otherwise): GpMpcController(process_group=...) over
NCCL -- sharded costs == unsharded costs per candidate, same arg-min and winner's trajectory on
every rank, same action from the sharded batched optimiser.

Tolerance 1e-9, not bit-for-bit: a candidate's N^2 partial sums meet in float64 reductions at L2 in scheduling order,
and the covariance sums cancel by ~1e8, so two evaluations of the SAME candidate differ by ~1e-10 on one GPU already
(tests/test_gpu_parity.py::test_batched_equals_looped_and_is_order_independent)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _controller(cfg, dev, group):
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.actions_config import ActionsConfig
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.config_classes.observation_config import ObservationConfig
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.config_classes.total_config import Config
    E, Na, H = cfg["E"], cfg["Na"], cfg["H"]
    r = cfg["reward"]
    config = Config(
        observation_config=ObservationConfig(obs_var_norm=[cfg["obs_var"]] * E),
        reward_config=RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                                   weight_state_terminal=list(r["weight_state_terminal"]),
                                   target_action_norm=list(r["target_action"]), weight_action=list(r["weight_action"]),
                                   exploration_factor=r["exploration_factor"]),
        actions_config=ActionsConfig(),
        controller_config=ControllerConfig(len_horizon=H, batched_candidates=24, batched_iters=6, batched_seed=11),
        model_config=ModelConfig(gp_init={"noise_covar.noise": list(cfg["noise"]),
                                          "base_kernel.lengthscale": [list(v) for v in cfg["lengthscale"]],
                                          "outputscale": list(cfg["outputscale"])},
                                 min_std_noise=1e-4, max_std_noise=1.0, min_outputscale=1e-6, max_outputscale=10.0,
                                 min_lengthscale=1e-3, max_lengthscale=1e3))
    c = GpMpcController(-np.ones(E), np.ones(E), -np.ones(Na), np.ones(Na), config, device=dev, process_group=group)
    c.transition_model.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    return c


def _outputs(c, cfg):
    a = torch.as_tensor(cfg["actions"].reshape(cfg["B"], -1))
    mu0, s0 = torch.as_tensor(cfg["mu0"]), torch.as_tensor(cfg["Sigma0"])
    costs, grads = c.compute_mean_lcb_trajectory_batch(a, mu0, s0)
    res = dict(costs=costs.cpu().numpy().copy(), grads=grads.cpu().numpy().copy(), best=c.best_candidate, shard=c.shard,
               mu=c.states_mu_pred.numpy().copy(), var=c.states_var_pred.numpy().copy())
    c.transition_model.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    act = c._get_optimal_actions_batched(mu0, s0)
    res.update(act=np.asarray(act).copy(), opt_cost=c.last_optim_cost)
    return res


def _worker(rank, world, port, distinct, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import tests.conftest  # noqa: F401  (package path)
    import torch.distributed as dist
    from oracle.workloads import make_workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = make_workload("C2", B=301, H=5, seed=61, distinct_lengthscales=distinct)
    res = _outputs(_controller(cfg, dev, True), cfg)
    if rank == 0:
        res["single"] = _outputs(_controller(cfg, dev, None), cfg)      # the same work on ONE GPU, no group
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("distinct", [False, True])
def test_sharded_costs_equal_single_gpu_costs_nccl(distinct):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, distinct, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    want = results[0]["single"]
    for rank in range(world):
        got = results[rank]
        lo, hi = got["shard"]
        assert (lo, hi) == ((0, 151) if rank == 0 else (151, 301))
        np.testing.assert_allclose(got["costs"], want["costs"], rtol=0, atol=1e-9)           # sharded == single GPU
        np.testing.assert_allclose(got["grads"], want["grads"][lo:hi], rtol=0, atol=1e-8)
        assert got["best"] == want["best"]
        np.testing.assert_allclose(got["mu"], want["mu"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(got["var"], want["var"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(got["act"], want["act"], rtol=0, atol=1e-6)
        assert abs(got["opt_cost"] - want["opt_cost"]) < 1e-8
    np.testing.assert_array_equal(results[0]["costs"], results[1]["costs"])                  # gathered: identical bits
    np.testing.assert_array_equal(results[0]["act"], results[1]["act"])
