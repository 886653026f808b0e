"""One-process-per-GPU helpers (torchrun): assemble the shards of a condensed distance vector.

The data path itself has no collective: rank r computes the row stripes dealt to shard r
(`b2d_shard_layout`) from the same inputs.  Only the finished distances travel - each rank's
compact slice list is gathered on one rank with `torch.distributed` (gloo on CPU tensors or
NCCL on GPU tensors) and scattered into the full SciPy-condensed vector.
"""

from __future__ import annotations

import numpy as np

from . import _capi


def scatter_compact(full, compact, layout):
    """Write a shard's compact result (stripes back to back) into the full condensed vector."""
    pos = 0
    for first, cnt in layout:
        full[first:first + cnt] = compact[pos:pos + cnt]
        pos += cnt
    return full


def gather_condensed(compact, n_samples, dst=0, group=None, device=None):
    """Gather every rank's compact shard on rank `dst`; returns the full vector there, None
    elsewhere.  `compact` is this rank's result in `b2d_problem_fetch_compact` order."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    total = n_samples * (n_samples - 1) // 2
    layouts = [_capi.shard_layout(n_samples, r, world) for r in range(world)]
    sizes = [t for _, t in layouts]
    assert len(compact) == sizes[rank], "compact shard has the wrong length for this rank"
    pad = max(sizes) if sizes else 0
    mine = torch.zeros(pad, dtype=torch.float64, device=device)
    mine[: sizes[rank]] = torch.as_tensor(np.ascontiguousarray(compact), device=device)
    bufs = [torch.zeros(pad, dtype=torch.float64, device=device) for _ in range(world)] \
        if rank == dst else None
    dist.gather(mine, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    full = np.empty(total, dtype=np.float64)
    for r in range(world):
        scatter_compact(full, bufs[r][: sizes[r]].cpu().numpy(), layouts[r][0])
    return full
