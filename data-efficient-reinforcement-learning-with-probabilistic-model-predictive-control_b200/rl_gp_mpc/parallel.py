"""Multi-GPU plumbing of the candidate batch.

Candidates never interact until the final arg-min, so the batch is cut into contiguous slices, one per rank;
the training block is replicated (every rank factorises the same (x, y, hyper-parameters)); the only collective
on the data path is ONE all-gather of the per-candidate costs, done in place on a buffer the rollout kernel writes
its slice of directly.  Gradients stay on the owning rank (a candidate's optimiser state lives with it); the winner's
trajectory (a few hundred doubles) is broadcast from its owner when the caller asks for the side-effect tensors.

`CandidateSharder` is what `GpMpcController(..., process_group=...)` uses; it only needs `torch.distributed`
(NCCL on the GPUs, gloo in the CPU tests) and tensors on the group's device."""
import torch


def shard_bounds(batch, world, rank):
    """Contiguous slice [lo, hi) of `batch` candidates owned by `rank`; `per` = padded slice length."""
    per = (batch + world - 1) // world
    lo = min(batch, rank * per)
    hi = min(batch, lo + per)
    return per, lo, hi


def allgather_costs(dist, costs_all, per, rank, group=None):
    """In-place all-gather: costs_all has world*per entries, rank r owns [r*per, (r+1)*per)."""
    dist.all_gather_into_tensor(costs_all, costs_all[rank * per:(rank + 1) * per], group=group)
    return costs_all


def global_argmin(costs_all, batch, per, world):
    """Index (into the unsharded batch) of the best finite cost after the all-gather."""
    idx = torch.arange(world * per, device=costs_all.device)
    r, k = idx // per, idx % per
    valid = (r * per + k < batch) & (r * per + k < (r + 1) * per) & torch.isfinite(costs_all)
    c = torch.where(valid, costs_all, torch.full_like(costs_all, float("inf")))
    j = int(torch.argmin(c).item())
    return j  # slices are contiguous: padded position == global candidate index


class CandidateSharder:
    """Contiguous sharding of a candidate batch over the ranks of a process group.

        sh = CandidateSharder(group)                  # group: a torch.distributed ProcessGroup, or None = WORLD
        lo, hi = sh.bounds(B)                         # this rank's candidates
        buf = sh.cost_buffer(B, device)               # (world * per,) -- the kernel writes buf[sh.local_view(B)]
        costs = sh.gather(buf, B)                     # ONE all-gather, in place; returns the (B,) view
        best = sh.argmin(buf, B)                      # same index on every rank
        owner = sh.owner(best, B)                     # rank that holds the winner's gradients / trajectory
        sh.broadcast(flat_tensor, owner)              # e.g. the winner's packed side-effect tensors
    """

    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("CandidateSharder needs an initialised torch.distributed process group")
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._bufs = {}

    def per(self, batch):
        return (batch + self.world - 1) // self.world

    def bounds(self, batch):
        _, lo, hi = shard_bounds(batch, self.world, self.rank)
        return lo, hi

    def cost_buffer(self, batch, device, dtype=torch.float64):
        """Reusable (world * per,) gather buffer; entries beyond the batch stay +inf."""
        key = (batch, str(device), dtype)
        buf = self._bufs.get(key)
        if buf is None:
            buf = torch.full((self.world * self.per(batch),), float("inf"), dtype=dtype, device=device)
            self._bufs[key] = buf
        return buf

    def local_view(self, buf, batch):
        """The slice of the gather buffer this rank's kernel writes (length hi - lo)."""
        lo, hi = self.bounds(batch)
        per = self.per(batch)
        return buf[self.rank * per: self.rank * per + (hi - lo)]

    def gather(self, buf, batch):
        """ONE in-place all-gather of the per-candidate costs; returns the (batch,) view every rank now holds."""
        allgather_costs(self.dist, buf, self.per(batch), self.rank, group=self.group)
        return buf[:batch]

    def argmin(self, buf, batch):
        return global_argmin(buf, batch, self.per(batch), self.world)

    def owner(self, index, batch):
        return min(self.world - 1, index // self.per(batch))

    def broadcast(self, tensor, owner):
        """Broadcast `tensor` (in place) from the group rank `owner`."""
        src = self.dist.get_global_rank(self.group, owner) if self.group is not None else owner
        self.dist.broadcast(tensor, src=src, group=self.group)
        return tensor
