"""Round-2 probe: the pipelined intersection kernel (b2d_isect.cuh) against the first-generation
kernels on BASELINE config 2 shapes - identical bits expected (integer sums), timings printed."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic

def run(key, n, tips, lam, variants, s0s=(0,)):
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    ids = tip_node[cols].astype(np.int32)
    ref = None
    out = {}
    with _capi.Problem.tree(parent, length, indptr, ids, vals, n) as prob:
        for s0 in s0s:
            _capi.set_option("sparse_rows_max_tips", s0)
            for (kern, grp) in variants:
                _capi.set_option("isect_kernel", kern)
                _capi.set_option("isect_group", grp)
                ts = []
                for rep in range(3):
                    prob.compute(key)
                    ts.append(prob.timings())
                t = ts[-1]
                got = prob.fetch(np.empty(n * (n - 1) // 2))
                same = None
                if s0 == s0s[0]:
                    if ref is None:
                        ref = got
                    same = bool(np.array_equal(ref, got))
                else:
                    same = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
                print(f"{key} n={n} tips={tips} S0={t['sparse_rows_max_tips']} kern={kern} G={grp}: tips_ms={t['tips_ms']:.2f} "
                      f"pair_ms={t['pair_ms']:.2f} compute_ms={t['compute_ms']:.2f} prep={t['skip_prep_ms']:.2f} "
                      f"matches={t['sparse_row_matches']:.3e} flagged={t['tips_flagged']} same={same}", flush=True)
                out[f"{key}_S{s0}_k{kern}_g{grp}"] = dict(tips_ms=t["tips_ms"], pair_ms=t["pair_ms"], compute_ms=t["compute_ms"],
                                                         matches=t["sparse_row_matches"], same=same, S0=t["sparse_rows_max_tips"])
    _capi.set_option("sparse_rows_max_tips", 0)
    _capi.set_option("isect_kernel", 1)
    _capi.set_option("isect_group", 4)
    return out

res = {}
V = [(0, 4), (1, 1), (1, 2), (1, 4), (1, 8)]
res.update(run("weighted_unifrac_normalized", 10000, 100000, 0.02, V))
res.update(run("unweighted_unifrac", 10000, 100000, 0.02, V))
res.update(run("weighted_unifrac_normalized", 10000, 100000, 0.02, [(1, 4)], s0s=(16, 24, 32, 48)))
res.update(run("unweighted_unifrac", 10000, 100000, 0.02, [(1, 4)], s0s=(8, 12, 16, 24, 32)))
json.dump(res, open("gpurun_out/r2_isect_probe.json", "w"), indent=1)
