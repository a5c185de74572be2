"""
Multi-GPU sharding of a likelihood batch (one process per GPU, `torch.distributed`).

The filters of one AMIS batch are independent (the reference evaluates them in a pure map,
/root/reference/bild/amis.py:735-739), so the batch is split into contiguous rank blocks with NO collective
inside the data path; the only exchange is one all-gather of the float64 logL vector per AMIS iteration
(NCCL over NVLink on GPUs, gloo in the CPU tests).  Every rank then holds the identical vector and runs the
identical fixed-order weight reduction, so the sampler state stays bit-identical across ranks and no
broadcast of proposals is needed (all ranks share the numpy seed).
"""
import numpy as np

__all__ = ["shard_bounds", "ShardedEvaluator"]


def shard_bounds(P, world):
    """Contiguous, balanced blocks: rank r owns [bounds[r], bounds[r+1])."""
    base, extra = divmod(int(P), int(world))
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:extra] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


class ShardedEvaluator:
    """
    Wraps a batch function ``f(ss, thetas) -> (P,) float64``: each rank evaluates its block, the blocks are
    all-gathered, every rank returns the full vector in the original order.

    Parameters
    ----------
    group : torch.distributed process group or None (default group)
    device : torch device the collective runs on ("cuda:<i>" for NCCL, "cpu" for gloo)
    """

    def __init__(self, group=None, device="cpu"):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist, self.group, self.device = dist, group, device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def __call__(self, f, ss, thetas):
        import torch
        P = len(ss)
        b = shard_bounds(P, self.world)
        lo, hi = int(b[self.rank]), int(b[self.rank + 1])
        mine = np.asarray(f(ss[lo:hi], thetas[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
        width = int(np.max(np.diff(b)))
        send = torch.zeros(width, dtype=torch.float64, device=self.device)
        send[:hi - lo] = torch.from_numpy(mine).to(self.device)
        recv = torch.empty(width * self.world, dtype=torch.float64, device=self.device)
        self.dist.all_gather_into_tensor(recv, send, group=self.group)
        recv = recv.cpu().numpy().reshape(self.world, width)
        return np.concatenate([recv[r, :int(b[r + 1] - b[r])] for r in range(self.world)])
