"""Round-2 probe: unweighted dense rows on the tensor cores (udense_tc) vs the LUT kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic
_capi.set_cache_limit(120 << 30)

def case(n, tips, lam, s0s, shard=0, n_shards=1, check=True):
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    ids = tip_node[cols].astype(np.int32)
    for s0 in s0s:
        _capi.set_option("sparse_rows_max_tips", s0)
        ref = None
        with _capi.Problem.tree(parent, length, indptr, ids, None, n, shard=shard, n_shards=n_shards) as prob:
            for tc in (0, 1):
                _capi.set_option("udense_tc", tc)
                for rep in range(3):
                    prob.compute("unweighted_unifrac")
                    t = prob.timings()
                rel = -1.0
                if check:
                    got, _ = prob.fetch_compact()
                    if ref is None:
                        ref = got
                    rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
                print(f"n={n} tips={tips} shard {shard}/{n_shards} S0={t['sparse_rows_max_tips']} tc={tc}: compute_ms={t['compute_ms']:.2f} pair_ms={t['pair_ms']:.2f} "
                      f"tips_ms={t['tips_ms']:.2f} prep={t['skip_prep_ms']:.2f} embed={t['embed_ms']:.2f} launches={t['pair_launches']} "
                      f"flagged={t['tips_flagged']} rel_vs_lut={rel:.2e}", flush=True)
    _capi.set_option("sparse_rows_max_tips", 0)
    _capi.set_option("udense_tc", 0)

which = sys.argv[1:] or ["small", "cfg2"]
if "small" in which: case(1000, 5000, 0.02, (0, 2))
if "cfg2" in which: case(10000, 100000, 0.02, (0, 4, 2))
if "cfg4" in which: case(100000, 500000, 0.004, (0, 20, 10), shard=0, n_shards=8, check=False)
