import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np
from skbio_b200 import _capi, synthetic
n, tips, lam = 10000, 100000, 0.02
parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
P = n * (n - 1) // 2
for s0 in (4, 6, 8, 12):
    _capi.set_option("sparse_rows_max_tips", s0)
    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], None, n) as prob:
        for rep in range(2):
            prob.compute("unweighted_unifrac")
        t = prob.timings()
        print(f"S0 {t['sparse_rows_max_tips']}: step {t['compute_ms']:.1f} ms embed {t['embed_ms']:.1f} pair {t['pair_ms']:.1f} tips {t['tips_ms']:.1f} prep {t['skip_prep_ms']:.1f} -> {P/(t['compute_ms']*1e-3):.3e} pairs/s", flush=True)
