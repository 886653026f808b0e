"""N > 1 host-side logic on CPU: world_size-2 gloo processes each own a shard of the pair
triangle (snake-dealt 128-sample row stripes), compute their slices, and rank 0 assembles the
full condensed vector.  The per-shard distances come from the oracle here (no GPU in this
test); on the GPU box the same layout is exercised by tests/test_gpu_parity.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, tips, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    from oracle import oracle_fast
    from skbio_b200 import _capi, distributed, synthetic

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=3)
        indptr, cols, vals = synthetic.poisson_csr(n, tips, 0.1, seed=4)
        dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
        full = oracle_fast.unifrac_all(dense, tip_node, parent, length)["weighted_unifrac_normalized"]
        layout, total = _capi.shard_layout(n, rank, world)
        compact = np.concatenate([full[a:a + c] for a, c in layout]) if len(layout) else np.zeros(0)
        assert len(compact) == total
        out = distributed.gather_condensed(compact, n, dst=0)
        if rank == 0:
            q.put(("ok", bool(np.array_equal(out, full)), total))
        else:
            q.put(("ok", out is None, total))
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e), 0))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_assemble():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, tips = 300, 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, tips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" and r[1] for r in res), res
    assert sum(r[2] for r in res) == n * (n - 1) // 2


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 129, 1000, 5000])
@pytest.mark.parametrize("shards", [1, 2, 3, 8])
def test_shard_layout_partitions_the_triangle(n, shards):
    from skbio_b200 import _capi

    total = n * (n - 1) // 2
    covered = np.zeros(total, dtype=np.int32)
    sizes = []
    for r in range(shards):
        layout, tot = _capi.shard_layout(n, r, shards)
        assert tot == sum(c for _, c in layout)
        sizes.append(tot)
        for a, c in layout:
            covered[a:a + c] += 1
    assert (covered == 1).all()
    if n >= 5000 and shards > 1:  # snake dealing balances the work
        assert max(sizes) / (sum(sizes) / shards) < 1.08


@pytest.mark.parametrize("n", [0, 1, 128, 129, 700, 10000, 28284])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_build_ranges_partition_the_sample_blocks(n, world):
    """The block ranges of the distributed build (b2d_build_range) are contiguous, disjoint, in rank
    order and cover every 128-sample block - what makes the exchange an in-place all-gather."""
    from skbio_b200 import _capi

    nb = (n + 127) // 128
    edge = 0
    for r in range(world):
        a, b = _capi.build_range(n, r, world)
        assert a == min(edge, nb) and a <= b <= nb
        edge = b
    assert edge == nb


def _blob_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    from skbio_b200 import distributed

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coll = distributed.TorchCollectives(None)
        mine = bytes([rank + 1]) * 24
        got = coll.allgather_host(mine)
        q.put(("ok", got == b"".join(bytes([r + 1]) * 24 for r in range(world)), coll.rank))
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e), rank))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_host_blobs_of_the_exchange():
    """The host half of the exchange (votes, entry counts, max U: TorchCollectives.allgather_host)
    between two gloo ranks; the device half needs GPUs (tests/test_gpu_parity.py runs it between
    host threads, bench.py under torchrun)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_blob_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" and r[1] for r in res), res


def test_thread_collectives_host_blobs():
    import threading

    from skbio_b200.distributed import ThreadCollectives

    world = 3
    coll = ThreadCollectives(world, timeout=30.0)
    got = [None] * world

    def work(r):
        c = coll.for_rank(r)
        for rep in range(3):
            got[r] = c.allgather_host(bytes([10 * rep + r]) * 4)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert all(g == b"".join(bytes([20 + r]) * 4 for r in range(world)) for g in got)
