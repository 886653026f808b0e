"""One shard of eight of BASELINE config 4 (100 000 x 500 000 tips, unweighted) computed twice - for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic

_capi.set_cache_limit(120 << 30)
parent, length, names, tip_node = synthetic.random_join_tree(500000, seed=1)
indptr, cols, vals = synthetic.poisson_csr(100000, 500000, 0.004, seed=0)
ids = tip_node[cols].astype(np.int32)
for k, v in (a.split("=") for a in sys.argv[1:]):
    _capi.set_option(k, int(v))
with _capi.Problem.tree(parent, length, indptr, ids, None, 100000, shard=0, n_shards=8) as prob:
    for rep in range(2):
        prob.compute("unweighted_unifrac")
        t = prob.timings()
    print({k: t[k] for k in ("tips_ms", "pair_ms", "compute_ms", "sparse_row_matches", "sparse_rows_max_tips", "isect_generation")})
