"""Config-2 problem with the lazy_proportions option off / on: phase timings (identical bits asserted)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic

n, tips, lam = 10000, 100000, 0.02
_capi.set_cache_limit(120 << 30)
parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
ids = tip_node[cols].astype(np.int32)
ref = None
for lazy in (0, 1, 0, 1):
    _capi.set_option("lazy_proportions", lazy)
    with _capi.Problem.tree(parent, length, indptr, ids, vals, n) as prob:
        for rep in range(4):
            prob.compute("weighted_unifrac_normalized")
            t = prob.timings()
        got = prob.fetch(np.empty(n * (n - 1) // 2))
    ref = got if ref is None else ref
    print(f"lazy={lazy}: compute_ms={t['compute_ms']:.2f} embed={t['embed_ms']:.2f} prep={t['skip_prep_ms']:.2f} "
          f"pair={t['pair_ms']:.2f} tips={t['tips_ms']:.2f} same={bool(np.array_equal(ref, got))}", flush=True)
_capi.set_option("lazy_proportions", 1)
