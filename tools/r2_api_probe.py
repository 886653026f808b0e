"""The public API path on config 2, several repetitions, with the host split and the engine's own phases."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import skbio_b200
from skbio_b200 import _capi, synthetic

n, tips = 10000, 100000
_capi.set_cache_limit(120 << 30)
parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
indptr, cols, vals = synthetic.poisson_csr(n, tips, 0.02, seed=0)
text = synthetic.to_newick(parent, length, names)
table = sp.csr_matrix((vals, cols, indptr), shape=(n, tips))
taxa = [f"T{i}" for i in range(tips)]
ids = [f"S{i}" for i in range(n)]
for rep in range(8):
    t0 = time.perf_counter()
    tree = skbio_b200.read_newick_flat(text)
    t1 = time.perf_counter()
    dm = skbio_b200.beta_diversity("weighted_unifrac", table, ids=ids, tree=tree, taxa=taxa, condensed=True, normalized=True)
    t2 = time.perf_counter()
    pr = skbio_b200.last_call_profile()
    print(f"rep{rep}: total {t2-t0:.4f} newick {t1-t0:.4f} " + " ".join(f"{k}={v:.4f}" for k, v in pr.items()), flush=True)
    del dm
# the same engine calls by hand
h = skbio_b200._driver._unifrac_host("weighted_unifrac_normalized", table, taxa, tree, True)
out = np.empty(n * (n - 1) // 2)
for rep in range(5):
    t0 = time.perf_counter()
    prob = _capi.Problem.tree(h["parent"], h["length"], h["indptr"], h["ids"], h["vals"], h["n"], node_of_col=h["node_of_col"])
    t1 = time.perf_counter(); prob.compute("weighted_unifrac_normalized"); t2 = time.perf_counter()
    prob.fetch(out); t3 = time.perf_counter(); tm = prob.timings(); prob.destroy(); t4 = time.perf_counter()
    print(f"hand{rep}: create {t1-t0:.4f} (upload {tm['upload_ms']:.1f} ms) compute {t2-t1:.4f} fetch {t3-t2:.4f} destroy {t4-t3:.4f}", flush=True)
