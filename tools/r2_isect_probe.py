"""Round-2 probe: the pipelined intersection kernel (b2d_isect.cuh) against the first-generation
kernels on BASELINE config 2 / 3 shapes - identical bits expected (integer sums), timings printed."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic

_capi.set_cache_limit(120 << 30)
res = {}

def tree_case(key, n, tips, lam, variants, s0s=(0,)):
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    ids = tip_node[cols].astype(np.int32)
    ref = None
    for s0 in s0s:
        _capi.set_option("sparse_rows_max_tips", s0)
        with _capi.Problem.tree(parent, length, indptr, ids, vals, n) as prob:
            for (kern, grp) in variants:
                _capi.set_option("isect_kernel", kern)
                _capi.set_option("isect_group", grp)
                for rep in range(3):
                    prob.compute(key)
                    t = prob.timings()
                got = prob.fetch(np.empty(n * (n - 1) // 2))
                if ref is None:
                    ref = got
                same = bool(np.array_equal(ref, got))
                rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
                print(f"{key} n={n} tips={tips} S0={t['sparse_rows_max_tips']} kern={kern} G={grp}: tips_ms={t['tips_ms']:.2f} "
                      f"pair_ms={t['pair_ms']:.2f} compute_ms={t['compute_ms']:.2f} prep={t['skip_prep_ms']:.2f} embed={t['embed_ms']:.2f} "
                      f"matches={t['sparse_row_matches']:.3e} entries={t['sparse_row_entries']:.3e} flagged={t['tips_flagged']} same={same} rel={rel:.1e}", flush=True)
                res[f"{key}_S{s0}_k{kern}_g{grp}"] = dict(tips_ms=t["tips_ms"], pair_ms=t["pair_ms"], compute_ms=t["compute_ms"],
                                                         matches=t["sparse_row_matches"], same=same, rel=rel, S0=t["sparse_rows_max_tips"])
    _capi.set_option("sparse_rows_max_tips", 0)
    _capi.set_option("isect_kernel", -1)
    _capi.set_option("isect_group", 0)

def feature_case(n, f, lam, variants):
    indptr, cols, vals = synthetic.poisson_csr(n, f, lam, seed=0)
    ref = {}
    with _capi.Problem.features(indptr, cols, vals.astype(np.float64), n, f) as prob:
        for (kern, grp) in variants:
            _capi.set_option("isect_kernel", kern)
            _capi.set_option("isect_group", grp)
            for key in ("braycurtis", "jaccard"):
                for rep in range(3):
                    prob.compute(key)
                    t = prob.timings()
                got, _ = prob.fetch_compact()
                chk = float(np.nansum(got[:1 << 22]))
                same = ref.setdefault(key, chk) == chk
                print(f"{key} n={n} f={f} kern={kern} G={grp}: pair_ms={t['pair_ms']:.2f} compute_ms={t['compute_ms']:.2f} same={same}", flush=True)
                res[f"{key}_k{kern}_g{grp}"] = dict(pair_ms=t["pair_ms"], same=same)
                del got
    _capi.set_option("isect_kernel", -1)
    _capi.set_option("isect_group", 0)

which = sys.argv[1:] or ["w", "u", "f", "ws", "us"]
V = [(0, 4), (1, 2), (1, 4), (1, 8)]
if "w" in which: tree_case("weighted_unifrac_normalized", 10000, 100000, 0.02, V)
if "u" in which: tree_case("unweighted_unifrac", 10000, 100000, 0.02, [(0, 4), (1, 1), (1, 2), (1, 4)])
if "f" in which: feature_case(50000, 50000, 0.01, [(0, 4), (1, 1), (1, 2), (1, 4)])
if "f1" in which: feature_case(50000, 50000, 0.01, [(1, 1)])
if "ws" in which: tree_case("weighted_unifrac_normalized", 10000, 100000, 0.02, [(1, 4)], s0s=(12, 16, 24, 32, 48))
if "ws2" in which: tree_case("weighted_unifrac_normalized", 10000, 100000, 0.02, [(-1, 0)], s0s=(14, 16, 18, 20, 22))
if "u4" in which:   # one shard of eight of BASELINE config 4
    parent, length, names, tip_node = synthetic.random_join_tree(500000, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(100000, 500000, 0.004, seed=0)
    ids = tip_node[cols].astype(np.int32)
    with _capi.Problem.tree(parent, length, indptr, ids, None, 100000, shard=0, n_shards=8) as prob:
        for (kern, grp) in [(0, 4), (1, 1), (1, 2)]:
            _capi.set_option("isect_kernel", kern)
            _capi.set_option("isect_group", grp)
            for rep in range(2):
                prob.compute("unweighted_unifrac")
                t = prob.timings()
            print(f"cfg4 shard 0/8 kern={kern} G={grp}: tips_ms={t['tips_ms']:.1f} pair_ms={t['pair_ms']:.1f} compute_ms={t['compute_ms']:.1f} "
                  f"prep={t['skip_prep_ms']:.1f} embed={t['embed_ms']:.1f} S0={t['sparse_rows_max_tips']} matches={t['sparse_row_matches']:.3e} "
                  f"entries={t['sparse_row_entries']:.3e}", flush=True)
    _capi.set_option("isect_kernel", -1)
    _capi.set_option("isect_group", 0)
if "u4s" in which:   # sparse-row threshold sweep on one shard of eight of BASELINE config 4
    parent, length, names, tip_node = synthetic.random_join_tree(500000, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(100000, 500000, 0.004, seed=0)
    ids = tip_node[cols].astype(np.int32)
    for s0 in (0,):   # the threshold is fixed when a problem first builds its sparse rows
        _capi.set_option("sparse_rows_max_tips", s0)
        with _capi.Problem.tree(parent, length, indptr, ids, None, 100000, shard=0, n_shards=8) as prob:
            for rep in range(2):
                prob.compute("unweighted_unifrac")
                t = prob.timings()
            print(f"cfg4 shard 0/8 S0={t['sparse_rows_max_tips']}: tips_ms={t['tips_ms']:.1f} pair_ms={t['pair_ms']:.1f} compute_ms={t['compute_ms']:.1f} "
                  f"prep={t['skip_prep_ms']:.1f} embed={t['embed_ms']:.1f} matches={t['sparse_row_matches']:.3e} "
                  f"entries={t['sparse_row_entries']:.3e}", flush=True)
    _capi.set_option("sparse_rows_max_tips", 0)
if "us" in which: tree_case("unweighted_unifrac", 10000, 100000, 0.02, [(1, 4)], s0s=(6, 8, 12, 16, 24, 32))
if "us2" in which: tree_case("unweighted_unifrac", 10000, 100000, 0.02, [(-1, 0)], s0s=(6, 7, 8, 9, 10, 12))
json.dump(res, open("gpurun_out/r2_isect_probe.json", "w"), indent=1)
