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
