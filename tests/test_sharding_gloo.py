"""CPU, world_size 2 over gloo: the N>1 plumbing (slice bounds, in-place cost all-gather, global arg-min)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl_gp_mpc.parallel import allgather_costs, global_argmin, shard_bounds


def test_shard_bounds_cover_the_batch_exactly():
    for batch in (1, 7, 8, 8192, 65536 + 3):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for rank in range(world):
                per, lo, hi = shard_bounds(batch, world, rank)
                assert 0 <= lo <= hi <= batch and hi - lo <= per
                seen.extend(range(lo, hi))
            assert seen == list(range(batch))


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = np.random.default_rng(5).uniform(0, 1, size=batch)           # the "whole job" every rank could compute
    costs[3] = np.nan
    per, lo, hi = shard_bounds(batch, world, rank)
    buf = torch.full((world * per,), -1.0, dtype=torch.float64)
    buf[rank * per: rank * per + (hi - lo)] = torch.as_tensor(costs[lo:hi])   # what the rollout kernel writes
    allgather_costs(dist, buf, per, rank)
    best = global_argmin(buf, batch, per, world)
    q.put((rank, buf[:batch].numpy().copy(), best))
    dist.destroy_process_group()


def test_allgather_of_costs_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    batch, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    costs = np.random.default_rng(5).uniform(0, 1, size=batch)
    costs[3] = np.nan
    want_best = int(np.nanargmin(costs))
    for rank, got, best in results:
        np.testing.assert_array_equal(got, costs)       # every rank holds every candidate's cost, bit-exact
        assert best == want_best


# ------------------------------------------------------------------------------------------------------------------
# GpMpcController(process_group=...) over gloo with a stand-in engine (CPU): the controller-level sharding -- slice, score,
# ONE all-gather, common arg-min, winner's trajectory broadcast -- against the same controller without a group.
# ------------------------------------------------------------------------------------------------------------------
_H, _NA, _E = 6, 1, 3


class _FakeEngine:
    """Deterministic analytic objective in place of the CUDA engine (device = cpu); honours `out` like the real one."""
    device = torch.device("cpu")

    def rollout(self, actions_mpc, obs_mu, obs_var, H_, iter_ctrl=0, limit_action_change=False, max_change=None,
                action_prev=None, need_grad=True, need_traj=True, out=None):
        a = torch.as_tensor(actions_mpc, dtype=torch.float64).reshape(-1, H_ * _NA)
        B = a.shape[0]
        target = torch.linspace(0.2, 0.8, H_ * _NA, dtype=torch.float64)
        w = torch.linspace(1.0, 2.0, H_ * _NA, dtype=torch.float64)
        o = out if out is not None else {}
        cost = (w * (a - target) ** 2).sum(1) + 0.125
        if "cost" in o:
            o["cost"].copy_(cost)
        else:
            o["cost"] = cost
        if need_grad:
            o["grad"] = 2 * w * (a - target)
        if need_traj:
            o["states_mu_pred"] = a[:, :1, None].expand(B, H_ + 1, _E).clone()
            o["states_var_pred"] = a[:, 1:2, None, None].expand(B, H_ + 1, _E, _E).clone()
            o["rewards_trajectory"] = -cost[:, None].expand(B, H_ + 1).clone()
            o["rewards_traj_var"] = a[:, 2:3].expand(B, H_ + 1).clone()
        return o


def _make_controller(group):
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.total_config import Config
    cfg = Config(controller_config=ControllerConfig(len_horizon=_H, batched_candidates=10, batched_iters=8, batched_seed=7))
    c = GpMpcController(-np.ones(_E), np.ones(_E), -np.ones(_NA), np.ones(_NA), cfg, process_group=group)
    c.transition_model._engine = _FakeEngine()
    c._cost_bound, c._cost_key_values = True, c._cost_fingerprint(cfg.reward)
    return c


def _controller_outputs(c):
    a = torch.as_tensor(np.random.default_rng(3).uniform(0, 1, size=(11, _H * _NA)))
    costs, grads = c.compute_mean_lcb_trajectory_batch(a, torch.zeros(_E, dtype=torch.float64), torch.eye(_E, dtype=torch.float64))
    res = dict(costs=costs.clone().numpy(), grads=grads.clone().numpy(), best=c.best_candidate, shard=c.shard,
               mu=c.states_mu_pred.numpy().copy(), var=c.states_var_pred.numpy().copy(), lcb=float(c.cost_traj_mean_lcb))
    act = c._get_optimal_actions_batched(torch.zeros(_E, dtype=torch.float64), torch.eye(_E, dtype=torch.float64))
    res.update(act=np.asarray(act).copy(), opt_cost=c.last_optim_cost, opt_costs=c.batched_costs.numpy().copy(),
               opt_lcb=float(c.cost_traj_mean_lcb))
    return res


def _controller_worker(rank, world, port, q):
    import sys
    for p in (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tests.conftest  # noqa: F401  (package path)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _controller_outputs(_make_controller(True))
    q.put((rank, res))
    dist.destroy_process_group()


def test_controller_shards_candidates_over_the_process_group_gloo():
    want = _controller_outputs(_make_controller(None))          # single process, no group
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_controller_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank in range(world):
        got = results[rank]
        lo, hi = got["shard"]
        assert (lo, hi) == ((0, 6) if rank == 0 else (6, 11))
        np.testing.assert_array_equal(got["costs"], want["costs"])           # every rank: every candidate's cost
        np.testing.assert_array_equal(got["grads"], want["grads"][lo:hi])    # gradients: the rank's own rows
        assert got["best"] == want["best"]
        np.testing.assert_array_equal(got["mu"], want["mu"])                 # winner's trajectory on every rank
        np.testing.assert_array_equal(got["var"], want["var"])
        assert got["lcb"] == want["lcb"]
        # batched optimiser: same restarts (shared seed), sliced over the ranks, one gather at the end
        np.testing.assert_allclose(got["opt_costs"], want["opt_costs"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(got["act"], want["act"], rtol=0, atol=1e-9)
        assert abs(got["opt_cost"] - want["opt_cost"]) < 1e-12 and abs(got["opt_lcb"] - want["opt_lcb"]) < 1e-12
    np.testing.assert_array_equal(results[0]["act"], results[1]["act"])      # both ranks apply the SAME action
