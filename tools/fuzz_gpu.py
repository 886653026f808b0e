"""Randomised parity sweep on the GPU (development aid): weighted UniFrac / dense Bray-Curtis through
the C ABI with random shapes, shards, node passes and launch chunks vs the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skbio_b200 import _capi, synthetic
from oracle import oracle_fast


def rel_err(a, b):
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    d[(a == b) | (np.isnan(a) & np.isnan(b))] = 0.0
    return float(np.max(d)) if len(d) else 0.0


def main(n_cases=40, seed=0, verbose=True):
    rng = np.random.default_rng(seed)
    worst = 0.0
    failures = []
    for case in range(n_cases):
        n = int(rng.choice([2, 3, 31, 33, 64, 97, 128, 129, 200, 257, 400, 700]))
        tips = int(rng.choice([2, 5, 63, 130, 500, 1500, 4000]))
        lam = float(rng.choice([0.002, 0.01, 0.05, 0.3, 1.0]))
        cat = bool(rng.integers(0, 2))
        gen = synthetic.caterpillar_tree if cat else synthetic.random_join_tree
        parent, length, names, tip_node = gen(tips, seed=int(rng.integers(1 << 30)))
        if rng.integers(0, 2):
            length = length * rng.uniform(1e-3, 1e3, size=len(length))
        indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=int(rng.integers(1 << 30)))
        dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
        exp = oracle_fast.unifrac_all(dense, tip_node, parent, length)
        n_shards = int(rng.choice([1, 1, 2, 3]))
        chunk = int(rng.choice([0, 128, 256, 1024]))
        rows_res = int(rng.choice([0, 128, 384, 1024]))
        max_tips = int(rng.choice([0, 0, 1, 3, 1 << 20]))   # sparse-row threshold (1 << 20: every row)
        _capi.set_option("sparse_rows_max_tips", max_tips)
        _capi.set_option("chunk_nodes", chunk)
        _capi.set_option("emb_budget_bytes", ((n + 127) // 128) * 128 * 8 * rows_res)
        total = n * (n - 1) // 2
        errs = {}
        for key in ("weighted_unifrac", "weighted_unifrac_normalized", "unweighted_unifrac"):
            out = np.full(total, -7.0)
            for sh in range(n_shards):
                with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n, shard=sh, n_shards=n_shards) as prob:
                    prob.compute(key)
                    prob.fetch(out)
            errs[key] = rel_err(out, exp[key])
        _capi.set_option("sparse_features", 0)
        with _capi.Problem.features(indptr, cols, vals.astype(np.float64), n, tips) as prob:
            prob.compute("braycurtis")
            got = prob.fetch(np.empty(total))
        _capi.set_option("sparse_features", -1)
        bc_ok = np.array_equal(got, exp["braycurtis"], equal_nan=True)
        w = max(errs.values())
        worst = max(worst, w)
        flag = "" if (w <= 1e-12 and bc_ok) else "  <-- FAIL"
        if flag:
            failures.append((case, n, tips, lam, cat, n_shards, chunk, rows_res, errs, bc_ok))
        if verbose:
            print(f"case {case}: n={n} tips={tips} lam={lam} cat={cat} shards={n_shards} chunk={chunk} rows={rows_res} maxtips={max_tips} "
                  f"wu {errs['weighted_unifrac']:.2e} wun {errs['weighted_unifrac_normalized']:.2e} "
                  f"uu {errs['unweighted_unifrac']:.2e} bc_exact {bc_ok}{flag}", flush=True)
    _capi.set_option("chunk_nodes", 0)
    _capi.set_option("emb_budget_bytes", 0)
    _capi.set_option("sparse_rows_max_tips", 0)
    if verbose:
        print("worst relative error", worst, "failures", len(failures))
    return worst, failures


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
