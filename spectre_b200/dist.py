"""Multi-rank glue for the sharded MSM: one process per GPU, point-range sharding, no data-path
collective -- only the 96-byte partial sums are all-gathered and folded on the host (EC addition is not an NCCL
reduction op). Backend-agnostic (`nccl` on the GPU box, `gloo` in the CPU tests)."""
import numpy as np

from . import halo2


def shard_range(n, rank, world):
    """Contiguous point range [lo, hi) owned by `rank` (same split the in-process multi-device context uses)."""
    return n * rank // world, n * (rank + 1) // world


def fold_partials(partials, world, device=None, group=None):
    """partials: (count, 12) uint64 Jacobian partial sums of this rank -> (count, 12) folded over all ranks."""
    import torch
    import torch.distributed as dist
    partials = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, 12)
    if world == 1:
        return partials
    t = torch.from_numpy(partials.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    allp = torch.stack(out).cpu().numpy().view(np.uint64)          # (world, count, 12)
    return np.stack([halo2.g1_sum(allp[:, i, :]) for i in range(allp.shape[1])])
