"""Multi-GPU plumbing of the candidate batch.

Candidates never interact until the final arg-min, so the batch is cut into contiguous slices, one per rank;
the training block is replicated (every rank factorises the same (x, y, hyper-parameters)); the only
collective is ONE all-gather of the per-candidate costs, done in place on a buffer the rollout kernel writes
its slice of directly.  Gradients stay on the owning rank."""


def shard_bounds(batch, world, rank):
    """Contiguous slice [lo, hi) of `batch` candidates owned by `rank`; `per` = padded slice length."""
    per = (batch + world - 1) // world
    lo = min(batch, rank * per)
    hi = min(batch, lo + per)
    return per, lo, hi


def allgather_costs(dist, costs_all, per, rank):
    """In-place all-gather: costs_all has world*per entries, rank r owns [r*per, (r+1)*per)."""
    dist.all_gather_into_tensor(costs_all, costs_all[rank * per:(rank + 1) * per])
    return costs_all


def global_argmin(costs_all, batch, per, world):
    """Index (into the unsharded batch) of the best finite cost after the all-gather."""
    import torch
    idx = torch.arange(world * per, device=costs_all.device)
    r, k = idx // per, idx % per
    valid = (r * per + k < batch) & (r * per + k < (r + 1) * per) & torch.isfinite(costs_all)
    c = torch.where(valid, costs_all, torch.full_like(costs_all, float("inf")))
    j = int(torch.argmin(c).item())
    return j  # slices are contiguous: padded position == global candidate index
