"""Multi-GPU plumbing of the two paths that shard: contiguous frame blocks (no collective) and the
row-sharded descriptor database whose per-shard top-2 results are all-gathered and merged exactly.
One process per GPU; `torch.distributed` (NCCL on the GPUs, gloo in the CPU tests) is only the transport."""
import numpy as np

from ._lib import TOP2_DTYPE
from .matcher import top2_merge


def shard_range(total, rank, world):
    """Contiguous block [r0, r1) of `total` units (frames or DB rows) owned by `rank`."""
    per = (total + world - 1) // world
    return min(rank * per, total), min((rank + 1) * per, total)


def allgather_merge_top2(local_top2, device=None, group=None):
    """local_top2: TOP2_DTYPE[nq] of this rank's shard (global row indices).  Every rank returns the merged
    TOP2_DTYPE[nq]: best = min distance (lowest global index on ties), second = second smallest of the union
    (ORBmatcher.cc:217-226 update rule applied to the concatenated database)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = np.ascontiguousarray(local_top2, TOP2_DTYPE)
    if world == 1:
        return top2_merge(local.reshape(1, -1))
    t = torch.from_numpy(local.view(np.uint8).reshape(len(local), TOP2_DTYPE.itemsize).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=torch.uint8, device=t.device)   # concatenated along dim 0
    dist.all_gather_into_tensor(out, t, group=group)
    parts = out.cpu().numpy().view(TOP2_DTYPE).reshape(world, len(local))
    return top2_merge(parts)
