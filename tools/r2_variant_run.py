"""Config-2 problem computed three times with a chosen intersection-kernel variant (for ncu captures).
usage: r2_variant_run.py <metric> <isect_kernel> <isect_group> [S0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic

key, kern, grp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
s0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n, tips, lam = 10000, 100000, 0.02
_capi.set_cache_limit(120 << 30)
parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
ids = tip_node[cols].astype(np.int32)
_capi.set_option("isect_kernel", kern)
_capi.set_option("isect_group", grp)
_capi.set_option("sparse_rows_max_tips", s0)
with _capi.Problem.tree(parent, length, indptr, ids, vals, n) as prob:
    for rep in range(3):
        prob.compute(key)
        t = prob.timings()
    print({k: t[k] for k in ("tips_ms", "pair_ms", "compute_ms", "sparse_row_matches")})
