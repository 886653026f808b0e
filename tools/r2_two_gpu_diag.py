"""Diagnostic: thread-collectives distributed build on devices (0, 0) and (0, 1) against one device."""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic
from skbio_b200.distributed import ThreadCollectives

n, tips = 900, 2400
parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=97)
indptr, cols, vals = synthetic.poisson_csr(n, tips, 0.03, seed=98)
ids = tip_node[cols].astype(np.int32)

def run(devs, key, exchange=True):
    world = len(devs)
    coll = ThreadCollectives(world, timeout=60.0)
    out = np.full(n * (n - 1) // 2, -1.0)
    errs, tms = [], [None] * world
    def work(r):
        try:
            with _capi.Problem.tree(parent, length, indptr, ids, vals, n, device=devs[r], shard=r, n_shards=world) as prob:
                if exchange and world > 1:
                    prob.set_collectives(r, world, coll.for_rank(r))
                prob.compute(key)
                tms[r] = prob.timings()
                prob.fetch(out)
        except BaseException as e:
            errs.append(e); coll.abort()
    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs, errs
    return out, tms

iu = np.triu_indices(n, 1)
for key in ("weighted_unifrac_normalized", "unweighted_unifrac"):
    base, _ = run([0], key)
    for devs, ex in (([0, 0], True), ([0, 1], False), ([0, 1], True), ([1, 0], True), ([1], True)):
        got, tms = run(devs, key, ex)
        bad = np.flatnonzero(got != base)
        msg = f"{key} devices={devs} exchange={ex}: mismatches {len(bad)}"
        if len(bad):
            bi, bj = iu[0][bad] // 128, iu[1][bad] // 128
            msg += f" row blocks {sorted(set(bi.tolist()))} col blocks {sorted(set(bj.tolist()))} max rel {np.max(np.abs(got[bad]-base[bad])/np.abs(base[bad])):.2e}"
        print(msg, [int(t['exchange_bytes']) for t in tms], flush=True)
