"""Quick on-GPU timing probe (development aid, not the benchmark)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic

def run(n, tips, lam, keys, feat_keys=()):
    t0 = time.time()
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    print(f"gen n={n} tips={tips} nnz={len(cols)} in {time.time()-t0:.1f}s", flush=True)
    P = n * (n - 1) // 2
    m = len(parent)
    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
        for key in keys:
            for rep in range(2):
                t0 = time.time(); prob.compute(key); wall = time.time() - t0
            t = prob.timings()
            np_per_s = P * m / (t["pair_ms"] * 1e-3)
            print(f"{key}: wall {wall*1e3:.1f} ms embed {t['embed_ms']:.2f} ms pair {t['pair_ms']:.2f} ms "
                  f"launches {t['pair_launches']} passes {t['passes']} -> {P/(t['compute_ms']*1e-3):.3e} pairs/s, "
                  f"{np_per_s:.3e} node-pairs/s = {np_per_s/148/1.9e9:.1f} np/clk/SM@1.9GHz", flush=True)
    if feat_keys:
        with _capi.Problem.features(indptr, cols, vals.astype(np.float64), n, tips) as prob:
            for key in feat_keys:
                for rep in range(2):
                    prob.compute(key)
                t = prob.timings()
                print(f"{key}: embed {t['embed_ms']:.2f} ms pair {t['pair_ms']:.2f} ms -> {P/(t['compute_ms']*1e-3):.3e} pairs/s", flush=True)

if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("e2e", "sparse", "permanova", "e2ef", "skip", "skip1")):
    print("dfma peak inst/s", _capi.microbench(0), "copy B/s", _capi.microbench(1), "lds64 gathers/s", _capi.microbench(2), flush=True)
    run(4096, 20000, 0.02, ["weighted_unifrac_normalized", "unweighted_unifrac"], ["braycurtis", "jaccard"])
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        run(10000, 100000, 0.02, ["weighted_unifrac_normalized", "unweighted_unifrac"])


def e2e_breakdown(n=10000, tips=100000, lam=0.02, key="weighted_unifrac_normalized"):
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    ids = tip_node[cols].astype(np.int32)
    P = n * (n - 1) // 2
    out = _capi.PinnedBuffer(P)
    for rep in range(12):
        t0 = time.perf_counter()
        prob = _capi.Problem.tree(parent, length, indptr, ids, vals, n)
        t1 = time.perf_counter()
        prob.compute(key)
        t2 = time.perf_counter()
        prob.fetch(out.array)
        t3 = time.perf_counter()
        tm = prob.timings()
        prob.destroy()
        t4 = time.perf_counter()
        print(f"e2e rep{rep}: create {1e3*(t1-t0):.1f} ms (upload {tm['upload_ms']:.1f}) compute {1e3*(t2-t1):.1f} ms "
              f"(dev {tm['compute_ms']:.1f}) fetch {1e3*(t3-t2):.1f} ms (dev {tm['fetch_ms']:.1f}) destroy {1e3*(t4-t3):.1f} ms "
              f"total {1e3*(t4-t0):.1f} ms", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "e2e":
    e2e_breakdown()


def sparse_probe(n=20000, f=50000, lam=0.01):
    indptr, cols, vals = synthetic.poisson_csr(n, f, lam, seed=0)
    P = n * (n - 1) // 2
    for mode in (1, 0):
        _capi.set_option("sparse_features", mode)
        with _capi.Problem.features(indptr, cols, vals.astype(np.float64), n, f) as prob:
            for key in ("braycurtis", "jaccard"):
                for rep in range(2):
                    t0 = time.perf_counter(); prob.compute(key); wall = time.perf_counter() - t0
                t = prob.timings()
                print(f"sparse={mode} {key}: wall {wall*1e3:.1f} ms embed {t['embed_ms']:.2f} pair {t['pair_ms']:.2f} ms -> {P/(t['compute_ms']*1e-3):.3e} pairs/s", flush=True)
    _capi.set_option("sparse_features", -1)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sparse":
    sparse_probe()


def permanova_probe(n=10000, perms=999):
    import skbio_b200
    from skbio_b200._dm import SimpleDistanceMatrix
    rng = np.random.default_rng(0)
    vec = rng.random(n * (n - 1) // 2)
    dm = SimpleDistanceMatrix(vec, [str(i) for i in range(n)], condensed=True)
    grouping = rng.integers(0, 4, size=n)
    for rep in range(2):
        t0 = time.perf_counter()
        res = skbio_b200.permanova(dm, grouping, permutations=perms, seed=1)
        dt = time.perf_counter() - t0
    print(f"permanova n={n} perms={perms}: {dt:.3f} s  F={res['test statistic']:.6f} p={res['p-value']}", flush=True)
    # CPU cost of ONE permutation with the reference's algorithm restated in NumPy
    t0 = time.perf_counter()
    iu = np.triu_indices(n, k=1)
    same = grouping[iu[0]] == grouping[iu[1]]
    sw = (vec[same] ** 2 / np.bincount(grouping)[grouping[iu[0]][same]]).sum()
    print(f"numpy one permutation: {time.perf_counter()-t0:.2f} s", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "permanova":
    permanova_probe()


def e2e_features(n=10000, f=50000, lam=0.01):
    indptr, cols, vals = synthetic.poisson_csr(n, f, lam, seed=0)
    P = n * (n - 1) // 2
    out = _capi.PinnedBuffer(P)
    for key, v in (("braycurtis", vals.astype(np.float64)), ("jaccard", vals.astype(np.float64)), ("braycurtis", vals.astype(np.float64))):
        for rep in range(2):
            t0 = time.perf_counter()
            prob = _capi.Problem.features(indptr, cols, v, n, f)
            t1 = time.perf_counter()
            prob.compute(key)
            t2 = time.perf_counter()
            prob.fetch(out.array)
            t3 = time.perf_counter()
            tm = prob.timings()
            prob.destroy()
            t4 = time.perf_counter()
            print(f"{key} rep{rep}: create {1e3*(t1-t0):.1f} compute {1e3*(t2-t1):.1f} (dev {tm['compute_ms']:.1f}, resident {tm['nodes_resident']}) fetch {1e3*(t3-t2):.1f} destroy {1e3*(t4-t3):.1f} ms", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "e2ef":
    e2e_features()


SKIPS = (0, 1)


def skip_probe(n=10000, tips=100000, lam=0.02):
    """weighted kernel with and without zero-block skipping (same problem, same box)"""
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    P = n * (n - 1) // 2
    res = {}
    for skip in SKIPS:
        _capi.set_option("weighted_skip", skip)
        with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
            for rep in range(2):
                prob.compute("weighted_unifrac_normalized")
            t = prob.timings()
            res[skip] = prob.fetch(np.empty(P))
            frac = t["visits_executed"] / t["visits_dense"] if t["visits_dense"] else 1.0
            print(f"n={n} tips={tips} lam={lam} skip={skip}: embed {t['embed_ms']:.2f} ms pair {t['pair_ms']:.2f} ms "
                  f"(prep {t['skip_prep_ms']:.2f} tips {t['tips_ms']:.2f} flagged {t['tips_flagged']} S0 {t['sparse_rows_max_tips']} entries {t['sparse_row_entries']}) launches {t['pair_launches']} visits {frac:.3f} of dense -> {P/(t['compute_ms']*1e-3):.3e} pairs/s", flush=True)
    _capi.set_option("weighted_skip", 1)
    for k in SKIPS[1:]:
        d = np.abs(res[0] - res[k]) / np.maximum(np.abs(res[0]), 1e-300)
        print(f"max rel diff skip={k} vs dense:", float(d.max()), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "skip":
    if len(sys.argv) > 2:
        SKIPS = tuple(int(x) for x in sys.argv[2].split(","))
    if len(sys.argv) > 3:
        _capi.set_option("sparse_rows_max_tips", int(sys.argv[3]))
        print("sparse_rows_max_tips", sys.argv[3])
    skip_probe(4096, 20000, 0.02)
    skip_probe()
    skip_probe(10000, 100000, 0.005)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "skip1":
    # one compute of the skipping path, for an ncu launch list
    n, tips, lam = 4096, 20000, 0.02
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
        prob.compute("weighted_unifrac_normalized")
        print(prob.timings())
