"""Multi-GPU helpers: the exchange step of the distributed build, and assembling the shards of a
condensed distance vector.

Rank r computes the row stripes dealt to shard r (`b2d_shard_layout`); its tiles need the
per-block products (sparse-row entries, packed dense rows, bitmaps, per-sample sums) of EVERY
column block.  With collectives installed (`Problem.set_collectives`) each rank builds those
products for its own contiguous block range only and the ranks all-gather them in place:

* `TorchCollectives` - one process per GPU under torchrun: `torch.distributed` (NCCL over NVLink /
  NVSwitch) on tensors that alias the library's device buffers;
* `ThreadCollectives` - one host thread per GPU inside one process
  (`beta_diversity(..., devices=[...])`): a barrier and peer copies.

The finished distances travel separately: each rank's compact slice list is gathered on one rank
(`gather_condensed`) or written by every shard into disjoint slices of one host vector.
"""

from __future__ import annotations

import threading

import numpy as np

from . import _capi


class _DeviceBytes:
    """`nbytes` of device memory at `ptr` for torch.as_tensor (CUDA array interface, no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class TorchCollectives:
    """The two collectives of the distributed build on a `torch.distributed` process group.

    NCCL group: device buffers are all-gathered in place (`all_gather_into_tensor` when every rank
    contributes the same number of bytes, grouped point-to-point transfers otherwise); the small
    host blobs ride through a device tensor.  A gloo group carries the host blobs only (CPU tests)."""

    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", int(device)) if device is not None else None
        self.on_device = dist.get_backend(group) == "nccl"
        self.calls = 0

    def install(self, problem):
        problem.set_collectives(self.rank, self.world, self)
        return problem

    def allgather_host(self, mine):
        torch, dist = self.torch, self.dist
        dev = self.device if self.on_device else "cpu"
        src = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
        out = torch.empty(len(mine) * self.world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, src, group=self.group)
        return out.cpu().numpy().tobytes()

    def allgatherv_device(self, bases, offsets, counts):
        torch, dist = self.torch, self.dist
        self.calls += 1
        pending, p2p = [], []
        with torch.cuda.device(self.device):
            for base, off, cnt in zip(bases, offsets, counts):
                lo = min(o for o, c in zip(off, cnt) if c) if any(cnt) else 0
                hi = max(o + c for o, c in zip(off, cnt) if c) if any(cnt) else 0
                if hi == lo:
                    continue
                whole = torch.as_tensor(_DeviceBytes(base + lo, hi - lo), device=self.device)
                views = [whole[off[r] - lo: off[r] - lo + cnt[r]] for r in range(self.world)]
                even = len(set(cnt)) == 1 and all(off[r] == lo + r * cnt[0] for r in range(self.world))
                if even:
                    pending.append(dist.all_gather_into_tensor(whole, views[self.rank], group=self.group, async_op=True))
                    continue
                # ranks contribute different byte counts (the last block range is shorter, entry counts
                # differ): every slice goes straight to its final place on every peer, all transfers of
                # all buffers in ONE NCCL group (full-duplex over NVLink / NVSwitch, nothing staged)
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    peer = dist.get_global_rank(self.group, r) if self.group is not None else r
                    if cnt[self.rank]:
                        p2p.append(dist.P2POp(dist.isend, views[self.rank], peer, self.group))
                    if cnt[r]:
                        p2p.append(dist.P2POp(dist.irecv, views[r], peer, self.group))
            if p2p:
                pending.extend(dist.batch_isend_irecv(p2p))
            for w in pending:
                w.wait()
            torch.cuda.current_stream(self.device).synchronize()


class ThreadCollectives:
    """The same two collectives between host threads of one process, one per shard: the blobs
    through a list, the device buffers by peer copies (`b2d_device_copy`).  `for_rank(r)` gives
    the object to install on shard r's problem; `abort()` releases ranks waiting for a failed one."""

    def __init__(self, world, timeout=600.0):
        self.world = int(world)
        self.barrier = threading.Barrier(self.world, timeout=timeout)
        self.blobs = [b""] * self.world
        self.bases = [None] * self.world

    def abort(self):
        self.barrier.abort()

    def for_rank(self, rank):
        return _ThreadRank(self, int(rank))


class _ThreadRank:
    def __init__(self, shared, rank):
        self.s, self.rank = shared, rank

    def allgather_host(self, mine):
        s = self.s
        s.blobs[self.rank] = bytes(mine)
        s.barrier.wait()
        out = b"".join(s.blobs)
        s.barrier.wait()          # nobody overwrites its blob before everybody has read all of them
        return out

    def allgatherv_device(self, bases, offsets, counts):
        s = self.s
        s.bases[self.rank] = list(bases)
        s.barrier.wait()
        for k, (off, cnt) in enumerate(zip(offsets, counts)):
            for r in range(s.world):
                if r != self.rank and cnt[r]:
                    _capi.device_copy(bases[k] + off[r], s.bases[r][k] + off[r], cnt[r])
        s.barrier.wait()          # every copy out of this rank's buffers is complete


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
