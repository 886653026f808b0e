"""In-process sharding: beta_diversity(..., devices=[0, 1]) on config 2 against one device (timing + bit equality)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import skbio_b200
from skbio_b200 import _capi, synthetic

n, tips = 10000, 100000
_capi.set_cache_limit(120 << 30)
parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
indptr, cols, vals = synthetic.poisson_csr(n, tips, 0.02, seed=0)
tree = skbio_b200.FlatTree(parent, length, names)
table = sp.csr_matrix((vals, cols, indptr), shape=(n, tips))
taxa = [f"T{i}" for i in range(tips)]
ref = None
for devs in ([0], [0, 1], [0], [0, 1]):
    t0 = time.perf_counter()
    dm = skbio_b200.beta_diversity("weighted_unifrac", table, tree=tree, taxa=taxa, condensed=True, normalized=True, devices=devs)
    dt = time.perf_counter() - t0
    vec = dm.condensed_form()
    ref = vec if ref is None else ref
    print(f"devices={devs}: {dt:.4f} s  same={bool(np.array_equal(ref, vec))}  {skbio_b200.last_call_profile().get('engine_s'):.4f} engine", flush=True)
