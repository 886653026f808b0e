"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI and the
Python host, against (a) the committed outputs of the reference itself and (b) the oracle on
seeded random inputs.

Tolerances (BASELINE.json north_star): integer/bit work is bit-exact (Jaccard, integer-count
Bray-Curtis, presence logic); UniFrac distances within 1e-12 relative in fp64.
"""
import os
import sys

import numpy as np
import pytest

from conftest import assert_close_rel, case_arrays, golden_cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIFRAC_RTOL = 1e-12


def _skbio_b200():
    import skbio_b200

    if not skbio_b200.available():
        pytest.fail("libskbb200.so missing or no CUDA device: GPU tests need the native engine")
    return skbio_b200


KEYS = [
    ("unweighted_unifrac", "unweighted_unifrac", {}),
    ("weighted_unifrac", "weighted_unifrac", {}),
    ("weighted_unifrac_normalized", "weighted_unifrac", {"normalized": True}),
    ("braycurtis", "braycurtis", {}),
    ("jaccard", "jaccard", {}),
]


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_golden_full_api(case):
    """Newick text -> tree object -> beta_diversity (engine) == reference output."""
    sb = _skbio_b200()
    counts, parent, length, names = case_arrays(case)
    tree = sb.read_newick(case["newick"])
    for key, metric, kw in KEYS:
        exp = np.array(case["results"][key], dtype=np.float64)
        kwargs = dict(kw)
        if "unifrac" in metric:
            kwargs.update(tree=tree, taxa=case["taxa"])
        dm = sb.beta_diversity(metric, counts, case["ids"], **kwargs)
        got = dm.condensed_form()
        assert list(dm.ids) == case["result_ids"]
        if "unifrac" in metric:  # same call with the native flat tree instead of node objects
            kwargs["tree"] = sb.read_newick_flat(case["newick"])
            dm2 = sb.beta_diversity(metric, counts, case["ids"], **kwargs)
            assert np.array_equal(dm2.condensed_form(), got, equal_nan=True)
        if key == "jaccard":
            assert np.array_equal(got, exp), f"{case['name']} jaccard not bit-exact"
        elif key == "braycurtis" and counts.dtype.kind in "iub":
            assert np.array_equal(got, exp, equal_nan=True), f"{case['name']} braycurtis not bit-exact"
        else:
            assert_close_rel(got, exp, UNIFRAC_RTOL, f"{case['name']} {key}")


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_faith_pd_matches_reference(case):
    sb = _skbio_b200()
    counts, parent, length, names = case_arrays(case)
    tree = sb.read_newick(case["newick"])
    got = sb.alpha_diversity("faith_pd", counts.astype(int) if counts.dtype != bool else counts,
                             case["ids"], tree=tree, taxa=case["taxa"])
    assert_close_rel(got.values, case["results"]["faith_pd"], UNIFRAC_RTOL, case["name"] + " faith_pd")


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_phydiv_matches_reference(case):
    """alpha_diversity('phydiv') for rooted/unrooted x unweighted/weighted/theta=0.5."""
    sb = _skbio_b200()
    counts, parent, length, names = case_arrays(case)
    tree = sb.read_newick_flat(case["newick"])
    icounts = counts.astype(int) if counts.dtype != bool else counts
    for rooted in (True, False):
        for weight in (False, True, 0.5):
            exp = case["results"][f"phydiv_r{int(rooted)}_w{float(weight)}"]
            got = sb.alpha_diversity("phydiv", icounts, case["ids"], tree=tree, taxa=case["taxa"],
                                     rooted=rooted, weight=weight, validate=False)
            assert_close_rel(got.values, exp, 1e-12, f"{case['name']} phydiv r={rooted} w={weight}")


def test_qiime_tiny_test(qiime_tt):
    sb = _skbio_b200()
    counts = np.array(qiime_tt["counts"])
    tree = sb.read_newick(qiime_tt["newick"])
    n = counts.shape[0]
    iu = np.triu_indices(n, k=1)
    for key, metric, kw in KEYS[:3]:
        dm = sb.beta_diversity(metric, counts, qiime_tt["ids"], tree=tree, taxa=qiime_tt["taxa"], **kw)
        got = dm.condensed_form()
        assert_close_rel(got, qiime_tt["reference_run"][key], UNIFRAC_RTOL, key)
        # QIIME 1.9.1 files carry 12 significant digits
        exp_q = np.array(qiime_tt["expected"][key])[iu]
        assert np.allclose(got, exp_q, rtol=0, atol=5e-12)


def _random_problem(n, tips, lam, seed, caterpillar=False):
    from skbio_b200 import synthetic

    gen = synthetic.caterpillar_tree if caterpillar else synthetic.random_join_tree
    parent, length, names, tip_node = gen(tips, seed=seed + 1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=seed)
    return parent, length, names, tip_node, indptr, cols, vals


@pytest.mark.parametrize("n,tips,lam,cat", [(300, 700, 0.05, False), (129, 1500, 0.01, True),
                                            (257, 64, 0.8, False)])
def test_random_vs_oracle_c_abi(n, tips, lam, cat):
    """Flat arrays straight into the C ABI vs the NumPy oracle (dense restatement)."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, lam, 3, cat)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    node_ids = tip_node[cols]
    exp = oracle_fast.unifrac_all(dense, tip_node, parent, length)
    with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
        for key in ("unweighted_unifrac", "weighted_unifrac", "weighted_unifrac_normalized"):
            prob.compute(key)
            got = prob.fetch(np.empty(n * (n - 1) // 2))
            assert_close_rel(got, exp[key], UNIFRAC_RTOL, key)
            if key == "unweighted_unifrac":
                assert_close_rel(prob.sample_vector(), exp["faith_pd"], UNIFRAC_RTOL, "faith_pd")
    with _capi.Problem.features(indptr, cols, vals.astype(np.float64), n, tips) as prob:
        prob.compute("braycurtis")
        got = prob.fetch(np.empty(n * (n - 1) // 2))
        assert np.array_equal(got, exp["braycurtis"], equal_nan=True)
        prob.compute("jaccard")
        got = prob.fetch(np.empty(n * (n - 1) // 2))
        assert np.array_equal(got, exp["jaccard"])


def test_plain_and_bulk_copy_kernels_agree_bitwise():
    _skbio_b200()
    from skbio_b200 import _capi

    n, tips = 200, 900
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.05, 9)
    node_ids = tip_node[cols]
    res = {}
    _capi.set_option("weighted_skip", 0)   # same node set and order in both kernels
    try:
        for plain in (0, 1):
            _capi.set_option("plain_kernels", plain)
            with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
                prob.compute("weighted_unifrac_normalized")
                res[plain] = prob.fetch(np.empty(n * (n - 1) // 2))
    finally:
        _capi.set_option("plain_kernels", 0)
        _capi.set_option("weighted_skip", 1)
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("n,tips,lam,cat,clade", [(700, 500, 0.1, False, False), (300, 3000, 0.004, False, False),
                                                  (161, 2500, 0.01, True, False), (40, 64, 0.5, False, False),
                                                  (300, 4000, 0.05, False, True), (200, 3000, 0.05, True, True)])
def test_zero_block_skipping_matches_the_dense_kernel(n, tips, lam, cat, clade):
    """The default weighted kernel skips nodes at which a 32-sample group is all zero and adds
    their closed-form contribution afterwards; it must agree with the kernel that visits every
    node (1e-13: summation order only) and with the oracle, also with several node passes,
    small launch chunks and shards."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, lam, 21, caterpillar=cat)
    if clade:
        # the first 96 samples (three whole groups) see only the first 40 % of the tips, and two
        # samples are empty: whole clades / the whole tree without a group (long zero paths)
        keep = np.ones(len(cols), dtype=bool)
        for s in range(min(96, n)):
            sl = slice(indptr[s], indptr[s + 1])
            keep[sl] = cols[sl] < int(0.4 * tips)
        for s in (100, 101):
            keep[indptr[s]:indptr[s + 1]] = False
        counts = np.add.reduceat(keep, indptr[:-1]) * (np.diff(indptr) > 0)
        cols, vals = cols[keep], vals[keep]
        indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    node_ids = tip_node[cols]
    total = n * (n - 1) // 2
    res = {}
    try:
        for skip in (0, 1):
            _capi.set_option("weighted_skip", skip)
            with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
                for key in ("weighted_unifrac", "weighted_unifrac_normalized"):
                    prob.compute(key)
                    t = prob.timings()
                    if skip:
                        assert 0 < t["visits_executed"] <= t["visits_dense"]
                    else:
                        assert t["visits_dense"] == 0
                    res[(skip, key)] = prob.fetch(np.empty(total))
        _capi.set_option("chunk_nodes", 128)
        _capi.set_option("emb_budget_bytes", ((n + 127) // 128) * 128 * 8 * 256)  # 256 node rows resident
        out = np.full(total, -1.0)
        for shard in range(2):
            with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n, shard=shard, n_shards=2) as prob:
                prob.compute("weighted_unifrac")
                if len(parent) > 256:
                    assert prob.timings()["passes"] > 1
                prob.fetch(out)
        res["multi"] = out
    finally:
        _capi.set_option("weighted_skip", 1)
        _capi.set_option("chunk_nodes", 0)
        _capi.set_option("emb_budget_bytes", 0)
    for key in ("weighted_unifrac", "weighted_unifrac_normalized"):
        assert_close_rel(res[(1, key)], res[(0, key)], 1e-13, key + " skip vs dense")
    assert_close_rel(res["multi"], res[(0, "weighted_unifrac")], 1e-13, "skip multi-pass shards")
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    exp = oracle_fast.unifrac_all(dense, tip_node, parent, length)
    for key in ("weighted_unifrac", "weighted_unifrac_normalized"):
        assert_close_rel(res[(1, key)], exp[key], UNIFRAC_RTOL, key + " skip vs oracle")


def test_shards_and_passes_reassemble_the_same_vector():
    """Row-stripe shards (snake dealing) fill disjoint slices; several node passes (small
    embedding budget) and small launch chunks give the same distances."""
    _skbio_b200()
    from skbio_b200 import _capi

    n, tips = 700, 500
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.1, 5)
    node_ids = tip_node[cols]
    total = n * (n - 1) // 2
    for key in ("weighted_unifrac", "unweighted_unifrac"):
        with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
            prob.compute(key)
            full = prob.fetch(np.empty(total))
        out = np.full(total, -1.0)
        covered = 0
        for shard in range(3):
            with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n, shard=shard,
                                    n_shards=3) as prob:
                prob.compute(key)
                covered += prob.result_len()
                prob.fetch(out)
        assert covered == total
        assert np.array_equal(out, full)
    # multi-pass + small chunks: summation order changes, values agree to 1e-13
    _capi.set_option("chunk_nodes", 256)
    _capi.set_option("emb_budget_bytes", 6 * 128 * 8 * 384)  # 384 node rows resident
    try:
        with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
            prob.compute("weighted_unifrac")
            t = prob.timings()
            multi = prob.fetch(np.empty(total))
        assert t["passes"] > 1
    finally:
        _capi.set_option("chunk_nodes", 0)
        _capi.set_option("emb_budget_bytes", 0)
    with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
        prob.compute("weighted_unifrac")
        single = prob.fetch(np.empty(total))
    assert_close_rel(multi, single, 1e-13, "multi-pass")


def test_properties_at_scale():
    """Size-independent properties on a larger case: identical samples are at distance 0,
    unweighted/normalized values lie in [0, 1], symmetric under sample permutation."""
    _skbio_b200()
    from skbio_b200 import _capi

    n, tips = 1500, 4000
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.02, 21)
    # make sample 7 a copy of sample 3
    rows = [(cols[indptr[s]:indptr[s + 1]], vals[indptr[s]:indptr[s + 1]]) for s in range(n)]
    rows[7] = rows[3]
    perm = np.random.default_rng(0).permutation(n)

    def build(order):
        ip = np.zeros(n + 1, dtype=np.int64)
        cs, vs = [], []
        for k, s in enumerate(order):
            cs.append(rows[s][0]); vs.append(rows[s][1])
            ip[k + 1] = ip[k] + len(rows[s][0])
        return ip, np.concatenate(cs), np.concatenate(vs)

    def run(order, key):
        ip, cs, vs = build(order)
        with _capi.Problem.tree(parent, length, ip, tip_node[cs], vs, n) as prob:
            prob.compute(key)
            return prob.fetch(np.empty(n * (n - 1) // 2))

    from scipy.spatial.distance import squareform
    for key in ("unweighted_unifrac", "weighted_unifrac_normalized"):
        a = squareform(run(np.arange(n), key))
        b = squareform(run(perm, key))
        assert a[3, 7] == 0.0
        assert a.min() >= 0.0 and a.max() <= 1.0 + 1e-12
        inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
        # the fp64 summation order depends on where a pair sits in its tile, so a permuted
        # run agrees to rounding, not bitwise
        assert_close_rel(a, b[np.ix_(inv, inv)], 1e-13, key + " permutation")
        assert np.array_equal(a, a.T)


def test_engine_errors_are_loud():
    sb = _skbio_b200()
    from skbio_b200 import _capi

    with pytest.raises(_capi.EngineError):
        _capi.Problem.tree(np.array([0, -1], dtype=np.int32), np.zeros(2), np.array([0, 0]),
                           np.zeros(0, dtype=np.int32), None, 1)  # parent[0] == 0 is invalid
    with pytest.raises(_capi.EngineError):
        _capi.Problem.tree(np.array([1, -1], dtype=np.int32), np.zeros(2), np.array([0, 1]),
                           np.array([5], dtype=np.int32), None, 1)  # node id out of range


def test_partial_and_block_beta_diversity():
    """partial_beta_diversity computes exactly the requested pairs (others 0) and
    block_beta_diversity equals the full matrix (reference: test_block.py:31-39,
    test_driver.py:814-951)."""
    sb = _skbio_b200()
    from skbio_b200 import synthetic

    n, tips = 300, 400
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=8)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, 0.05, seed=9)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    tree = synthetic.to_nodes(parent, length, names)
    taxa = [f"T{i}" for i in range(tips)]
    ids = [f"S{i}" for i in range(n)]
    for metric, kw in (("unweighted_unifrac", {}), ("weighted_unifrac", {"normalized": True})):
        full = sb.beta_diversity(metric, dense, ids, tree=tree, taxa=taxa, **kw)
        fdata = np.asarray(full.data)
        rng = np.random.default_rng(1)
        pairs = set()
        while len(pairs) < 200:
            i, j = rng.integers(0, n, size=2)
            if i != j and (j, i) not in pairs:
                pairs.add((int(i), int(j)))
        pairs = sorted(pairs)
        part = sb.partial_beta_diversity(metric, dense, ids, [(ids[i], ids[j]) for i, j in pairs],
                                         tree=tree, taxa=taxa, **kw)
        pdata = np.asarray(part.data)
        mask = np.zeros((n, n), dtype=bool)
        for i, j in pairs:
            mask[i, j] = mask[j, i] = True
        assert_close_rel(pdata[mask], fdata[mask], 1e-13, metric + " partial")
        assert not pdata[~mask].any()
        blk = sb.block_beta_diversity(metric, dense, ids, tree=tree, taxa=taxa, k=64, **kw)
        assert np.array_equal(np.asarray(blk.data), fdata)
        # the reference's map/reduce structure over k x k blocks of ids
        blk2 = sb.block_beta_diversity(metric, dense, ids, tree=tree, taxa=taxa, k=100,
                                       map_f=lambda f, gen: [f(**kwargs) for kwargs in gen], **kw)
        assert_close_rel(np.asarray(blk2.data), fdata, 1e-13, metric + " block")


@pytest.mark.parametrize("n,f,lam", [(700, 9000, 0.01), (257, 5000, 0.02), (130, 300, 1.5)])
def test_sparse_and_dense_feature_paths_are_bit_identical(n, f, lam):
    """Bray-Curtis / Jaccard on integer tables: the sparse intersection kernels and the dense
    kernels give the same bits, and both equal SciPy's definitions (oracle)."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    indptr, cols, vals = synthetic.poisson_csr(n, f, lam, seed=17)
    # a few empty samples and a large count
    vals = vals.copy()
    if len(vals):
        vals[0] = 100000
    dense = synthetic.csr_to_dense(indptr, cols, vals, f)
    exp_bc, exp_j = oracle_fast.braycurtis(dense), oracle_fast.jaccard(dense)
    total = n * (n - 1) // 2
    res = {}
    try:
        for mode in (0, 1):
            _capi.set_option("sparse_features", mode)
            with _capi.Problem.features(indptr, cols, vals.astype(np.float64), n, f) as prob:
                for key in ("braycurtis", "jaccard"):
                    prob.compute(key)
                    t = prob.timings()
                    if mode == 0 or lam < 0.1:   # very dense buckets legitimately fall back
                        assert (t["nodes_resident"] == 0) == (mode == 1)
                    res[(mode, key)] = prob.fetch(np.empty(total))
    finally:
        _capi.set_option("sparse_features", -1)
    for key, exp in (("braycurtis", exp_bc), ("jaccard", exp_j)):
        assert np.array_equal(res[(0, key)], exp, equal_nan=True), key + " dense"
        assert np.array_equal(res[(1, key)], exp, equal_nan=True), key + " sparse"


def test_sparse_path_falls_back_for_non_integer_counts():
    _skbio_b200()
    from skbio_b200 import _capi, synthetic

    n, f = 200, 4000
    indptr, cols, vals = synthetic.poisson_csr(n, f, 0.01, seed=3)
    fv = vals.astype(np.float64) * 0.5
    dense = synthetic.csr_to_dense(indptr, cols, fv, f)
    from oracle import oracle_fast
    with _capi.Problem.features(indptr, cols, fv, n, f) as prob:
        prob.compute("braycurtis")
        assert prob.timings()["nodes_resident"] > 0      # dense fp64 path
        got = prob.fetch(np.empty(n * (n - 1) // 2))
    assert_close_rel(got, oracle_fast.braycurtis(dense), 1e-13, "float braycurtis")


def test_config1_full_size_through_the_api():
    """BASELINE config 1 in full (1,000 samples x 10,000 tips, 19,999 nodes, Poisson 0.1) through
    the public API with a tree OBJECT and shuffled taxa, every pair compared with the oracle."""
    import time

    sb = _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import synthetic

    n, tips = 1000, 10000
    parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, 0.1, seed=0)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    perm = np.random.default_rng(5).permutation(tips)          # taxa in shuffled order
    taxa = [f"T{i}" for i in perm]
    table = np.ascontiguousarray(dense[:, perm])
    tree = synthetic.to_nodes(parent, length, names)
    ids = [f"S{i}" for i in range(n)]
    exp = oracle_fast.unifrac_all(dense, tip_node, parent, length)
    timings = {}
    for key, metric, kw in KEYS:
        kwargs = dict(kw)
        if "unifrac" in metric:
            kwargs.update(tree=tree, taxa=taxa)
        t0 = time.perf_counter()
        dm = sb.beta_diversity(metric, table, ids, **kwargs)
        timings[key] = time.perf_counter() - t0
        got = dm.condensed_form()
        if key in ("jaccard", "braycurtis"):
            assert np.array_equal(got, exp[key], equal_nan=True), key
        else:
            assert_close_rel(got, exp[key], UNIFRAC_RTOL, "config 1 " + key)
    print("config-1 full API wall seconds:", {k: round(v, 3) for k, v in timings.items()})


@pytest.mark.parametrize("cat", [False, True])
def test_decomposed_and_sequential_walks_agree_bitwise(cat):
    """The subtree-decomposed weighted propagation (parallel small subtrees + short top walk)
    and the sequential per-lane walk produce identical distances; both match the oracle."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 130, 6000
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.02, 31, cat)
    node_ids = tip_node[cols]
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    exp = oracle_fast.unifrac_all(dense, tip_node, parent, length)["weighted_unifrac_normalized"]
    res = {}
    try:
        for mode in (1, 0):
            _capi.set_option("walk_decomposed", mode)
            with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
                prob.compute("weighted_unifrac_normalized")
                res[mode] = prob.fetch(np.empty(n * (n - 1) // 2))
    finally:
        _capi.set_option("walk_decomposed", 1)
    assert np.array_equal(res[0], res[1])
    assert_close_rel(res[1], exp, UNIFRAC_RTOL, "decomposed walk")


def _random_nested_tree(rng, n_tips, max_children=5, p_unary=0.1):
    """Random multifurcating tree as nested tuples (name, length, children); unary internal
    nodes, zero-length branches and a root length included."""
    nodes = [(f"t{i}", float(rng.choice([0.0, rng.uniform(0.01, 2.0)], p=[0.1, 0.9])), [])
             for i in range(n_tips)]
    k = 0
    while len(nodes) > 1:
        if rng.random() < p_unary:
            arity = 1
        else:
            arity = int(rng.integers(2, max_children + 1))
        arity = min(arity, len(nodes))
        idx = sorted(rng.choice(len(nodes), size=arity, replace=False), reverse=True)
        kids = [nodes.pop(i) for i in idx]
        length = float(rng.choice([0.0, rng.uniform(0.01, 2.0)], p=[0.1, 0.9]))
        nodes.append((f"n{k}" if rng.random() < 0.5 else None, length, kids))
        k += 1
        if arity == 1 and len(nodes) == 1:
            break
    name, _, kids = nodes[0]
    return (name, float(rng.choice([0.0, 0.3])) if rng.random() < 0.5 else None, kids)


def _nested_to_newick(node):
    name, length, kids = node
    s = ""
    if kids:
        s += "(" + ",".join(_nested_to_newick(c) for c in kids) + ")"
    if name is not None:
        s += "'" + name + "'"
    if length is not None:
        s += ":" + repr(length)
    return s


@pytest.mark.parametrize("seed", range(6))
def test_random_multifurcating_trees_full_api_vs_numpy_oracle(seed):
    """Random multifurcating trees with unary nodes, zero lengths and optional root length:
    Newick -> product flattener -> GPU, against the NumPy oracle fed by its OWN, separate
    implementation of the reference's id assignment."""
    sb = _skbio_b200()
    from oracle import oracle_np

    rng = np.random.default_rng(100 + seed)
    n_tips = int(rng.integers(5, 120))
    nested = _random_nested_tree(rng, n_tips)
    newick = _nested_to_newick(nested) + ";"
    parent, length, names, child_index = oracle_np.tree_arrays_from_nested(nested)
    n = int(rng.integers(2, 40))
    tips = [f"t{i}" for i in rng.permutation(n_tips)]
    taxa = tips[: max(1, int(n_tips * rng.uniform(0.5, 1.0)))]
    counts = rng.poisson(rng.uniform(0.2, 2.0), size=(n, len(taxa)))
    counts[rng.integers(0, n)] = 0
    tree = sb.read_newick(newick)
    p2, l2, n2 = sb.flatten_tree(tree)
    assert np.array_equal(p2, parent) and np.array_equal(l2, length) and list(n2) == list(names)
    for key, metric, kw in KEYS[:3]:
        exp = oracle_np.unifrac_condensed(metric, counts, taxa, parent, length, names,
                                          normalized=kw.get("normalized", False))
        with np.errstate(all="ignore"):
            # validate=False: the random root may have more than two children ("unrooted")
            got = sb.beta_diversity(metric, counts, tree=tree, taxa=taxa, validate=False,
                                    **kw).condensed_form()
        assert_close_rel(got, exp, UNIFRAC_RTOL, f"seed {seed} {key}")
    assert np.array_equal(sb.beta_diversity("jaccard", counts).condensed_form(),
                          oracle_np.jaccard_condensed(counts))
    assert np.array_equal(sb.beta_diversity("braycurtis", counts).condensed_form(),
                          oracle_np.braycurtis_condensed(counts), equal_nan=True)


def test_all_zero_table_and_degenerate_shapes():
    """An all-zero table: UniFrac variants and Jaccard give 0.0, Bray-Curtis NaN (SciPy's 0/0),
    exactly like the reference; 2-sample and 128/129-sample shapes hit the tile edges."""
    sb = _skbio_b200()
    from oracle import oracle_np

    tree = sb.read_newick("((a:1,b:2):1,(c:1,d:3):2)r;")
    taxa = list("abcd")
    zero = np.zeros((3, 4), dtype=int)
    assert not sb.beta_diversity("unweighted_unifrac", zero, tree=tree, taxa=taxa).condensed_form().any()
    assert not sb.beta_diversity("weighted_unifrac", zero, tree=tree, taxa=taxa).condensed_form().any()
    assert not sb.beta_diversity("weighted_unifrac", zero, tree=tree, taxa=taxa,
                                 normalized=True).condensed_form().any()
    assert not sb.beta_diversity("jaccard", zero).condensed_form().any()
    assert np.isnan(sb.beta_diversity("braycurtis", zero).condensed_form()).all()
    parent, length, names = sb.flatten_tree(tree)
    rng = np.random.default_rng(2)
    for n in (2, 128, 129):
        counts = rng.poisson(1.0, size=(n, 4))
        for metric, kw in (("unweighted_unifrac", {}), ("weighted_unifrac", {"normalized": True})):
            exp = oracle_np.unifrac_condensed(metric, counts, taxa, parent, length, names,
                                              normalized=kw.get("normalized", False))
            with np.errstate(all="ignore"):
                got = sb.beta_diversity(metric, counts, tree=tree, taxa=taxa, **kw).condensed_form()
            assert_close_rel(got, exp, UNIFRAC_RTOL, f"n={n} {metric}")


def test_permanova_matches_reference():
    """GPU PERMANOVA (one pass over all permutations) vs the reference's permanova on the same
    distance matrices, groupings and seeds: pseudo-F to 1e-12, p-value identical (the
    permutation stream is reproduced on the host with the same Generator calls)."""
    import json

    from conftest import GOLDEN

    sb = _skbio_b200()
    from skbio_b200._dm import SimpleDistanceMatrix

    with open(os.path.join(GOLDEN, "permanova.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        n = c["n"]
        dm = SimpleDistanceMatrix(np.array(c["condensed"]), [f"s{i}" for i in range(n)])
        res = sb.permanova(dm, c["grouping"], permutations=c["permutations"], seed=c["seed"])
        assert res["sample size"] == n and res["number of groups"] == c["num_groups"]
        assert abs(res["test statistic"] - c["stat"]) <= 1e-12 * abs(c["stat"])
        if c["permutations"] == 0:
            assert np.isnan(res["p-value"])
        else:
            assert res["p-value"] == c["p_value"]
    with pytest.raises(ValueError, match="All values in the grouping vector are the same"):
        sb.permanova(dm, ["a"] * n)
    with pytest.raises(ValueError, match="Grouping vector size must match"):
        sb.permanova(dm, ["a", "b"])
    with pytest.raises(TypeError, match="Input must be a DistanceMatrix"):
        sb.permanova(np.zeros((3, 3)), ["a", "b", "a"])


def test_reference_driver_known_answers():
    """The reference's own beta_diversity tests, on the engine
    (/root/reference/skbio/diversity/tests/test_driver.py:364-385, 625-713): hand-computed UniFrac
    values to 6 places, Bray-Curtis goldens, DataFrame index -> ids, counts vs presence/absence
    equality for qualitative metrics."""
    import pandas as pd

    sb = _skbio_b200()
    table1 = [[1, 5], [2, 3], [0, 1]]
    sids1 = list("ABC")
    tree1 = sb.read_newick("((O1:0.25, O2:0.50):0.25, O3:0.75)root;")
    oids1 = ["O1", "O2"]
    dm = sb.beta_diversity("unweighted_unifrac", table1, sids1, taxa=oids1, tree=tree1)
    assert dm.shape == (3, 3)
    np.testing.assert_almost_equal(np.asarray(dm.data), [[0, 0, 0.25], [0, 0, 0.25], [0.25, 0.25, 0]], 6)
    dm = sb.beta_diversity("weighted_unifrac", table1, sids1, taxa=oids1, tree=tree1)
    np.testing.assert_almost_equal(np.asarray(dm.data),
                                   [[0, 0.175, 0.12499999], [0.175, 0, 0.3], [0.12499999, 0.3, 0]], 6)
    dm = sb.beta_diversity("weighted_unifrac", table1, sids1, taxa=oids1, tree=tree1, normalized=True)
    np.testing.assert_almost_equal(np.asarray(dm.data),
                                   [[0, 0.128834, 0.085714], [0.128834, 0, 0.2142857], [0.085714, 0.2142857, 0]], 6)
    dm = sb.beta_diversity("braycurtis", table1, sids1)
    np.testing.assert_almost_equal(dm.condensed_form(), [0.27272727, 0.71428571, 0.66666667], 6)
    table2 = [[23, 64, 14, 0, 0, 3, 1], [0, 3, 35, 42, 0, 12, 1], [0, 5, 5, 0, 40, 40, 0],
              [44, 35, 9, 0, 1, 0, 0], [0, 2, 8, 0, 35, 45, 1], [0, 0, 25, 35, 0, 19, 0]]
    sids2 = list("ABCDEF")
    dm = sb.beta_diversity("braycurtis", table2, sids2)
    exp = [[0, 0.78787879, 0.86666667, 0.30927835, 0.85714286, 0.81521739],
           [0.78787879, 0, 0.78142077, 0.86813187, 0.75, 0.1627907],
           [0.86666667, 0.78142077, 0, 0.87709497, 0.09392265, 0.71597633],
           [0.30927835, 0.86813187, 0.87709497, 0, 0.87777778, 0.89285714],
           [0.85714286, 0.75, 0.09392265, 0.87777778, 0, 0.68235294],
           [0.81521739, 0.1627907, 0.71597633, 0.89285714, 0.68235294, 0]]
    np.testing.assert_almost_equal(np.asarray(dm.data), exp, 6)
    # DataFrame index becomes the ids (issue 1808)
    df = pd.DataFrame(table2, index=sids2)
    assert list(sb.beta_diversity("jaccard", df).ids) == sids2
    # qualitative metrics: counts and presence/absence give the same matrix (issue 1549)
    table3 = table2 + [[88, 31, 0, 5, 5, 5, 5], [44, 39, 0, 0, 0, 0, 0]]
    pa = np.asarray(table3) > 0
    a = sb.beta_diversity("jaccard", table3).condensed_form()
    b = sb.beta_diversity("jaccard", pa).condensed_form()
    assert np.array_equal(a, b)
    assert not np.array_equal(sb.beta_diversity("braycurtis", table3).condensed_form(),
                              sb.beta_diversity("braycurtis", pa.astype(int)).condensed_form())


def test_randomised_sweep_weighted_and_dense_braycurtis():
    """Random shapes (n = 2 .. 700, 2 .. 4000 tips, both tree kinds, scaled branch lengths), shards,
    node passes and launch chunks through the C ABI: weighted UniFrac within 1e-12 of the oracle,
    dense Bray-Curtis bit-exact (tools/fuzz_gpu.py, 16 cases)."""
    _skbio_b200()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_gpu

    worst, failures = fuzz_gpu.main(16, seed=5, verbose=False)
    assert not failures, failures
    assert worst <= UNIFRAC_RTOL


def test_sparse_tip_rows_agree_with_the_dense_kernel_and_fall_back_when_they_must():
    """Tip rows as a fixed-point sparse intersection (default) vs the dense kernel; the two
    fall-backs: a table too dense for the position buckets, and near-duplicate samples (where
    U_i + U_j - 2 * intersection would cancel): those pairs alone are recomputed in floating point."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    def run(parent, length, indptr, cols, vals, tip_node, n, key, tips_sparse):
        _capi.set_option("tips_sparse", tips_sparse)
        try:
            with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
                prob.compute(key)
                return prob.fetch(np.empty(n * (n - 1) // 2)), prob.timings()
        finally:
            _capi.set_option("tips_sparse", 1)

    # 1. ordinary sparse table: sparse tips used, no pair flagged, agrees with dense tips and oracle
    n, tips = 300, 3000
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.02, 31)
    exp = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr, cols, vals, tips), tip_node, parent, length)
    for key in ("weighted_unifrac", "weighted_unifrac_normalized"):
        a, ta = run(parent, length, indptr, cols, vals, tip_node, n, key, 1)
        b, tb = run(parent, length, indptr, cols, vals, tip_node, n, key, 0)
        assert ta["tips_ms"] > 0 and ta["tips_flagged"] == 0 and tb["tips_ms"] == 0
        assert_close_rel(a, b, 1e-13, key + " sparse vs dense tips")
        assert_close_rel(a, exp[key], UNIFRAC_RTOL, key + " sparse tips vs oracle")

    # 2. identical samples stay at exactly 0; a near-duplicate pair triggers the dense redo
    vals2, cols2, indptr2 = vals.copy(), cols.copy(), indptr.copy()
    s0, s1 = slice(indptr[0], indptr[1]), None
    # sample 1 := sample 0 (exact copy), sample 2 := sample 0 with one count changed by 1 in ~1e6
    rows = [(cols[s0], vals[s0] * 100000), (cols[s0], vals[s0] * 100000)]
    v3 = vals[s0] * 100000
    v3 = v3.copy()
    v3[0] += 1
    rows.append((cols[s0], v3))
    for s in range(3, n):
        rows.append((cols[indptr[s]:indptr[s + 1]], vals[indptr[s]:indptr[s + 1]]))
    cols2 = np.concatenate([r[0] for r in rows])
    vals2 = np.concatenate([r[1] for r in rows])
    indptr2 = np.concatenate([[0], np.cumsum([len(r[0]) for r in rows])]).astype(np.int64)
    exp2 = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr2, cols2, vals2, tips), tip_node, parent, length)
    got, t = run(parent, length, indptr2, cols2, vals2, tip_node, n, "weighted_unifrac", 1)
    assert t["tips_flagged"] > 0 and t["tips_ms"] > 0           # only the flagged pairs are redone (fp64)
    assert got[0] == 0.0                                          # pair (0, 1): identical samples
    assert_close_rel(got, exp2["weighted_unifrac"], UNIFRAC_RTOL, "near-duplicates")

    # 3. every sample holds every tip: position buckets overflow, the dense kernel takes the tips
    n3, tips3 = 260, 4500
    parent3, length3, names3, tip_node3, indptr3, cols3, vals3 = _random_problem(n3, tips3, 3.0, 32)
    exp3 = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr3, cols3, vals3, tips3), tip_node3, parent3, length3)
    got3, t3 = run(parent3, length3, indptr3, cols3, vals3, tip_node3, n3, "weighted_unifrac_normalized", 1)
    assert t3["tips_ms"] == 0 and t3["tips_flagged"] == 0
    assert_close_rel(got3, exp3["weighted_unifrac_normalized"], UNIFRAC_RTOL, "dense table")


@pytest.mark.parametrize("cat", [False, True])
def test_sparse_rows_threshold_does_not_change_the_result(cat):
    """Rows of nodes with up to k tips below go through the intersection kernel (k from the table's
    density by default): k = 1 (tips only), 2, 5, 40, automatic, and the all-dense kernel agree."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 390, 5000
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.01, 41, caterpillar=cat)
    exp = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr, cols, vals, tips), tip_node, parent, length)
    total = n * (n - 1) // 2
    res = {}
    try:
        for k in (0, 1, 2, 5, 40):
            _capi.set_option("sparse_rows_max_tips", k)
            with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
                prob.compute("weighted_unifrac_normalized")
                t = prob.timings()
                assert t["tips_ms"] > 0 and t["tips_flagged"] == 0
                assert t["sparse_rows_max_tips"] == (k if k else t["sparse_rows_max_tips"]) >= 1
                assert t["sparse_row_matches"] > 0
                res[k] = prob.fetch(np.empty(total))
        _capi.set_option("tips_sparse", 0)
        with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
            prob.compute("weighted_unifrac_normalized")
            res["dense"] = prob.fetch(np.empty(total))
    finally:
        _capi.set_option("sparse_rows_max_tips", 0)
        _capi.set_option("tips_sparse", 1)
    for k, v in res.items():
        assert_close_rel(v, exp["weighted_unifrac_normalized"], UNIFRAC_RTOL, f"sparse rows k={k}")
        assert_close_rel(v, res["dense"], 1e-13, f"sparse rows k={k} vs dense")


def test_one_problem_many_computes_in_any_order():
    """A problem keeps its buffers (embedding, sparse-row entries, packed rows, bitmaps) between
    computes; switching metrics back and forth must give what a fresh problem gives."""
    _skbio_b200()
    from skbio_b200 import _capi

    n, tips = 260, 2500
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.02, 51)
    node_ids = tip_node[cols]
    total = n * (n - 1) // 2
    fresh = {}
    for key in ("weighted_unifrac", "unweighted_unifrac", "weighted_unifrac_normalized"):
        with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
            prob.compute(key)
            fresh[key] = prob.fetch(np.empty(total))
    with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
        pd_fresh = prob.phydiv(True, 0.5)
    order = ["weighted_unifrac", "unweighted_unifrac", "weighted_unifrac_normalized", "weighted_unifrac",
             "unweighted_unifrac", "weighted_unifrac_normalized", "weighted_unifrac_normalized"]
    with _capi.Problem.tree(parent, length, indptr, node_ids, vals, n) as prob:
        for step, key in enumerate(order):
            prob.compute(key)
            got = prob.fetch(np.empty(total))
            assert np.array_equal(got, fresh[key]), (step, key)
            if step == 3:
                assert np.array_equal(prob.phydiv(True, 0.5), pd_fresh)


def test_negative_branch_lengths_keep_every_row_in_the_dense_kernel():
    """len * min(p_i, p_j) is only min(len p_i, len p_j) for len >= 0: a tree with a negative
    branch length must not take the sparse-row path, and still matches the oracle."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 150, 900
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.03, 61)
    length = length.copy()
    length[tip_node[:50]] *= -1.0
    exp = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr, cols, vals, tips), tip_node, parent, length)
    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
        prob.compute("weighted_unifrac")
        t = prob.timings()
        got = prob.fetch(np.empty(n * (n - 1) // 2))
    assert t["tips_ms"] == 0 and t["sparse_rows_max_tips"] == 0
    assert_close_rel(got, exp["weighted_unifrac"], 1e-9, "negative lengths")  # signed sums cancel


@pytest.mark.parametrize("cat", [False, True])
def test_unweighted_sparse_rows_agree_with_the_lut_kernel_and_fall_back(cat):
    """Unweighted UniFrac: sparse rows through the key-intersection kernel + packed dense stripes
    (default) vs every row in the look-up-table kernel, for several thresholds; near-duplicate
    samples make the engine redo the LUT way; Faith's PD is untouched."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 333, 4000
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.015, 71, caterpillar=cat)
    exp = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr, cols, vals, tips), tip_node, parent, length)
    total = n * (n - 1) // 2
    res = {}
    try:
        for k in (0, 1, 4, 50, 1 << 20):
            _capi.set_option("sparse_rows_max_tips", k)
            for shards in (1, 2):
                out = np.full(total, -1.0)
                for sh in range(shards):
                    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], None, n, shard=sh, n_shards=shards) as prob:
                        prob.compute("unweighted_unifrac")
                        t = prob.timings()
                        assert t["tips_ms"] > 0 and t["tips_flagged"] == 0 and t["sparse_row_matches"] > 0
                        prob.fetch(out)
                        if shards == 1:
                            assert_close_rel(prob.sample_vector(), exp["faith_pd"], UNIFRAC_RTOL, "faith_pd")
                res[(k, shards)] = out
        _capi.set_option("tips_sparse", 0)
        with _capi.Problem.tree(parent, length, indptr, tip_node[cols], None, n) as prob:
            prob.compute("unweighted_unifrac")
            assert prob.timings()["tips_ms"] == 0
            res["lut"] = prob.fetch(np.empty(total))
    finally:
        _capi.set_option("sparse_rows_max_tips", 0)
        _capi.set_option("tips_sparse", 1)
    for k, v in res.items():
        assert_close_rel(v, exp["unweighted_unifrac"], UNIFRAC_RTOL, f"unweighted sparse rows {k}")
        assert_close_rel(v, res["lut"], 1e-13, f"unweighted sparse rows {k} vs LUT")

    # an exact copy stays at exactly 0
    rows = [cols[indptr[s]:indptr[s + 1]] for s in range(n)]
    rows[1] = rows[0].copy()
    cols2 = np.concatenate(rows)
    indptr2 = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    dense2 = np.zeros((n, tips), dtype=np.int64)
    for s in range(n):
        dense2[s, rows[s]] = 1
    exp2 = oracle_fast.unifrac_all(dense2, tip_node, parent, length)
    with _capi.Problem.tree(parent, length, indptr2, tip_node[cols2], None, n) as prob:
        prob.compute("unweighted_unifrac")
        got = prob.fetch(np.empty(total))
    assert got[0] == 0.0
    assert_close_rel(got, exp2["unweighted_unifrac"], UNIFRAC_RTOL, "unweighted with a duplicate sample")

    # a near-duplicate (one extra taxon on a 1e-9 branch whose parent the sample already reaches):
    # the guard has the engine recompute that pair from the presence bits, every row
    in0 = np.zeros(tips, dtype=bool)
    in0[rows[0]] = True
    par_of_tip = parent[tip_node]
    extra = None
    for t in np.flatnonzero(~in0):
        sib = np.flatnonzero((par_of_tip == par_of_tip[t]) & in0)
        if len(sib):
            extra = int(t)
            break
    assert extra is not None
    length3 = length.copy()
    length3[tip_node[extra]] = 1e-9
    rows[2] = np.concatenate([rows[0], [extra]])
    cols3 = np.concatenate(rows)
    indptr3 = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    dense3 = np.zeros((n, tips), dtype=np.int64)
    for s in range(n):
        dense3[s, rows[s]] = 1
    exp3 = oracle_fast.unifrac_all(dense3, tip_node, parent, length3)
    with _capi.Problem.tree(parent, length3, indptr3, tip_node[cols3], None, n) as prob:
        prob.compute("unweighted_unifrac")
        t = prob.timings()
        got = prob.fetch(np.empty(total))
    assert t["tips_flagged"] > 0 and t["tips_ms"] > 0   # the flagged pairs alone are recomputed
    assert_close_rel(got, exp3["unweighted_unifrac"], UNIFRAC_RTOL, "unweighted near-duplicates")


@pytest.mark.parametrize("key", ["weighted_unifrac_normalized", "unweighted_unifrac"])
def test_pipelined_intersection_kernel_matches_the_first_generation_bit_for_bit(key):
    """b2d_isect.cuh (producer warp + bulk-copy ring, G lanes per row entry) against k_pair_tips /
    k_pair_utips: the sparse-row sums are integers, so every variant must give identical bits -
    single-batch buckets (sparse table) and buckets cut into several batches (denser table,
    several sample blocks), shards included."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    for n, tips, lam, seed in ((300, 3000, 0.02, 81), (520, 2600, 0.06, 82)):
        parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, lam, seed)
        exp = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr, cols, vals, tips), tip_node, parent, length)
        total = n * (n - 1) // 2
        res = {}
        try:
            for kern, grp in ((0, 4), (1, 1), (1, 2), (1, 4), (1, 8)):
                _capi.set_option("isect_kernel", kern)
                _capi.set_option("isect_group", grp)
                for shards in (1, 3):
                    out = np.full(total, -1.0)
                    for sh in range(shards):
                        with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n, shard=sh,
                                                n_shards=shards) as prob:
                            prob.compute(key)
                            t = prob.timings()
                            assert t["tips_ms"] > 0 and t["sparse_row_matches"] > 0
                            prob.fetch(out)
                    res[(kern, grp, shards)] = out
        finally:
            _capi.set_option("isect_kernel", -1)
            _capi.set_option("isect_group", 0)
        base = res[(0, 4, 1)]
        assert_close_rel(base, exp[key], UNIFRAC_RTOL, key + " first-generation kernel vs oracle")
        for k, v in res.items():
            assert np.array_equal(v, base), (key, n, k)


def test_flagged_pairs_are_recomputed_alone():
    """Three near-duplicate pairs among 1 500 samples: the guard flags exactly those, the intersection
    kernel still runs (no dense redo of the step), the result is within 1e-12 of the oracle, exact
    copies stay at exactly 0, and the step takes about as long as without them."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 1500, 6000
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.02, 91)
    rows = [(cols[indptr[s]:indptr[s + 1]].copy(), vals[indptr[s]:indptr[s + 1]].copy() * 1000000) for s in range(n)]

    def timed(rows_):
        c = np.concatenate([r[0] for r in rows_])
        v = np.concatenate([r[1] for r in rows_])
        ip = np.concatenate([[0], np.cumsum([len(r[0]) for r in rows_])]).astype(np.int64)
        with _capi.Problem.tree(parent, length, ip, tip_node[c], v, n) as prob:
            ms = []
            for _ in range(4):
                prob.compute("weighted_unifrac")
                ms.append(prob.timings()["compute_ms"])
            return prob.fetch(np.empty(n * (n - 1) // 2)), prob.timings(), min(ms[1:]), (ip, c, v)

    base, t0, ms0, _ = timed(rows)
    assert t0["tips_flagged"] == 0 and t0["tips_ms"] > 0
    dup = [list(r) for r in rows]
    for a, b in ((7, 700), (130, 131), (900, 1499)):   # three tiles, b := a with one count off by 1e-6
        c, v = rows[a][0].copy(), rows[a][1].copy()
        v[len(v) // 2] += 1
        dup[b] = [c, v]
    dup[3] = [rows[2][0].copy(), rows[2][1].copy()]   # an exact copy as well
    got, t1, ms1, (ip, c, v) = timed([tuple(r) for r in dup])
    assert t1["tips_flagged"] == 3 and t1["tips_ms"] > 0
    dense = synthetic.csr_to_dense(ip, c, v, tips)
    exp = oracle_fast.weighted(oracle_fast.counts_by_node(dense, tip_node, parent),
                               np.ascontiguousarray(length, dtype=np.float64), parent, False)
    assert_close_rel(got, exp, UNIFRAC_RTOL, "three flagged pairs")
    assert got[2 * n - (2 + 2) * (2 + 1) // 2 + 3] == 0.0   # pair (2, 3): exact copies
    assert ms1 <= 1.25 * ms0 + 0.5, (ms0, ms1)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_device_embedding_equals_vectorize_counts_and_tree(case):
    """K1 pinned directly (mirrors /root/reference/skbio/diversity/tests/test_util.py:27-34): the
    node-count matrix the device builds (CSR scatter + tree propagation) is, in reference id order,
    exactly what vectorize_counts_and_tree returned for the same inputs - with the sequential walk,
    with the subtree-decomposed walk, with the walk cut into node passes, and as presence bits."""
    sb = _skbio_b200()
    from oracle import oracle_np
    from skbio_b200 import _capi, _driver

    counts, parent, length, names = case_arrays(case)
    exp = np.array(case["counts_by_node"], dtype=np.int64)
    tree = sb.read_newick_flat(case["newick"])
    p, l, ip, ids, vals, n = _driver._unifrac_host_arrays("weighted_unifrac", counts, case["taxa"], tree, False)
    m = len(p)
    assert exp.shape == (n, m)
    with _capi.Problem.tree(p, l, ip, ids, vals, n) as prob:
        assert np.array_equal(prob.fetch_embedding(0, m), exp)
        assert np.array_equal(prob.fetch_embedding(1, m), exp)
        try:
            _capi.set_option("emb_budget_bytes", ((n + 127) // 128) * 128 * 8 * 128)   # one 128-node stripe per pass
            assert np.array_equal(prob.fetch_embedding(0, m), exp)
        finally:
            _capi.set_option("emb_budget_bytes", 0)
    # presence: the reference vectorises (counts > 0) for unweighted UniFrac (_driver.py:303-304)
    p, l, ip, ids, vals, n = _driver._unifrac_host_arrays("unweighted_unifrac", counts, case["taxa"], tree, False)
    with _capi.Problem.tree(p, l, ip, ids, None, n) as prob:
        bits = prob.fetch_embedding(2, m)
    pres = oracle_np.nodes_by_counts(np.atleast_2d(counts) > 0, case["taxa"], names,
                                     oracle_np.child_index_from_parent(parent))
    assert np.array_equal(bits, (np.asarray(pres).T > 0).astype(np.int64))


@pytest.mark.parametrize("cat", [False, True])
def test_device_embedding_random_trees_and_passes(cat):
    """The same on seeded random inputs (random-join and caterpillar trees, 300 samples over three
    sample blocks), whole tree resident and in several node passes, against the C oracle."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 300, 1500
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.03, 17, caterpillar=cat)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    exp = oracle_fast.counts_by_node(dense, tip_node, parent)
    m = len(parent)
    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
        assert np.array_equal(prob.fetch_embedding(0, m), exp)
        assert np.array_equal(prob.fetch_embedding(1, m), exp)
        try:
            _capi.set_option("emb_budget_bytes", 384 * 8 * 256)   # 256 node rows per pass
            assert np.array_equal(prob.fetch_embedding(0, m), exp)
        finally:
            _capi.set_option("emb_budget_bytes", 0)
        assert np.array_equal(prob.fetch_embedding(2, m), (exp > 0).astype(np.int64))


def test_devices_argument_runs_one_shard_per_host_thread():
    """beta_diversity(..., devices=[0, 0]): two shards, two host threads, (here) one device, each
    writing its own stripes of one condensed vector - identical to the single-shard call
    (tiles == full, /root/reference/skbio/diversity/_block.py:263-356)."""
    sb = _skbio_b200()
    from skbio_b200 import synthetic

    n, tips = 700, 1200
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.04, 23)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    tree = sb.FlatTree(parent, length, names)
    taxa = [f"T{i}" for i in range(tips)]
    for metric, kw in (("weighted_unifrac", {"normalized": True}), ("unweighted_unifrac", {}),
                       ("braycurtis", {}), ("jaccard", {})):
        if "unifrac" in metric:
            kw = dict(kw, tree=tree, taxa=taxa)
        one = sb.beta_diversity(metric, dense, **kw).condensed_form()
        for devs in ([0, 0], [0, 0, 0]):
            many = sb.beta_diversity(metric, dense, devices=devs, **kw).condensed_form()
            assert np.array_equal(one, many, equal_nan=True), (metric, devs)


def _run_shards_in_threads(world, make, key, computes=2, exchange=True):
    """One problem per shard, one host thread per shard (all on device 0), the distributed build's
    collectives between the threads: (full condensed vector, timings of every shard)."""
    import threading

    from skbio_b200.distributed import ThreadCollectives

    coll = ThreadCollectives(world, timeout=120.0)
    n = make(0, 1).n
    out = np.full(n * (n - 1) // 2, -1.0)
    tms, errors = [None] * world, []

    def work(r):
        try:
            with make(r, world) as prob:
                if exchange:
                    prob.set_collectives(r, world, coll.for_rank(r))
                for _ in range(computes):     # the second compute reuses every buffer of the first
                    prob.compute(key)
                tms[r] = prob.timings()
                prob.fetch(out)
        except BaseException as e:
            errors.append(e)
            coll.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return out, tms


@pytest.mark.parametrize("key", ["weighted_unifrac_normalized", "weighted_unifrac", "unweighted_unifrac"])
def test_distributed_build_gives_the_same_bits_as_the_replicated_one(key):
    """b2d_problem_set_collectives: every shard builds the embedding, sparse-row entries, packed
    dense rows and bitmaps of its own block range only and the shards all-gather them; the pair
    kernels then see exactly the arrays a replicated build produces, so the distances are
    bit-identical - block counts that divide evenly, that do not, and more shards than blocks
    (a shard with nothing to build votes no and everybody builds everything)."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    for n, tips, lam, seed in ((700, 1200, 0.04, 91), (530, 2600, 0.02, 92)):
        parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, lam, seed)
        exp = oracle_fast.unifrac_all(synthetic.csr_to_dense(indptr, cols, vals, tips), tip_node, parent, length)

        def make(r, w):
            return _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n, shard=r, n_shards=w)

        base, tm = _run_shards_in_threads(1, make, key)
        assert tm[0]["exchange_bytes"] == 0
        assert_close_rel(base, exp[key], UNIFRAC_RTOL, key + " single shard vs oracle")
        nb = (n + 127) // 128
        for world in (2, 3, 4, 8):
            got, tms = _run_shards_in_threads(world, make, key)
            assert np.array_equal(got, base), (key, n, world)
            per = (nb + world - 1) // world
            everyone_builds = per * (world - 1) < nb
            for t in tms:
                assert (t["exchange_bytes"] > 0) == everyone_builds, (key, n, world, t)
            rep, tms = _run_shards_in_threads(world, make, key, exchange=False)
            assert np.array_equal(rep, base)
            assert all(t["exchange_bytes"] == 0 for t in tms)
            # partial upload: every shard uploads the CSR rows of its own block range only, the first
            # compute gathers the others' device to device (also when the build itself is replicated)
            try:
                _capi.set_option("partial_upload", 1)
                got, tms2 = _run_shards_in_threads(world, make, key)
            finally:
                _capi.set_option("partial_upload", 0)
            assert np.array_equal(got, base), (key, n, world, "partial upload")
            full_bytes = (n + 1) * 8 + len(cols) * 12
            assert sum(t["csr_upload_bytes"] for t in tms2) == full_bytes + (world - 1) * (n + 1) * 8


def test_lazy_proportions_option_gives_identical_bits():
    """Option lazy_proportions: the consumers of the embedding divide count / total themselves instead of a
    separate pass over the whole embedding - the same correctly rounded division, so identical bits
    (near-duplicate pairs included: their recomputation reads the counts too)."""
    _skbio_b200()
    from skbio_b200 import _capi

    n, tips = 520, 2600
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.05, 95)
    res = {}
    try:
        for lazy in (0, 1):
            _capi.set_option("lazy_proportions", lazy)
            for key in ("weighted_unifrac_normalized", "weighted_unifrac"):
                with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
                    prob.compute(key)
                    assert prob.timings()["tips_ms"] > 0
                    res[(lazy, key)] = prob.fetch(np.empty(n * (n - 1) // 2))
    finally:
        _capi.set_option("lazy_proportions", 0)
    for key in ("weighted_unifrac_normalized", "weighted_unifrac"):
        assert np.array_equal(res[(0, key)], res[(1, key)]), key


def test_distributed_build_votes_no_when_a_compute_cannot_use_it():
    """Collectives installed, but the compute cannot be distributed: several node passes (the embedding
    does not fit the budget), Faith's PD (no pair work), a problem smaller than the number of shards.
    Every rank still takes part in the vote, the build is replicated, the bits are the same."""
    _skbio_b200()
    from skbio_b200 import _capi

    n, tips = 420, 1800
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.03, 96)

    def make(r, w):
        return _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n, shard=r, n_shards=w)

    base, _ = _run_shards_in_threads(1, make, "weighted_unifrac_normalized", computes=1)
    try:
        _capi.set_option("emb_budget_bytes", 512 * 8 * 512)   # a few hundred node rows per pass
        multi, _ = _run_shards_in_threads(1, make, "weighted_unifrac_normalized", computes=1)
        got, tms = _run_shards_in_threads(2, make, "weighted_unifrac_normalized")
    finally:
        _capi.set_option("emb_budget_bytes", 0)
    assert np.array_equal(got, multi)
    assert_close_rel(got, base, 1e-13, "several passes vs one")
    assert all(t["exchange_bytes"] == 0 and t["passes"] > 1 for t in tms)
    # 130 samples = 2 blocks, 4 shards: shards 2 and 3 have nothing to build
    n2 = 130
    ip2, c2, v2 = indptr[: n2 + 1], cols[: indptr[n2]], vals[: indptr[n2]]

    def make2(r, w):
        return _capi.Problem.tree(parent, length, ip2, tip_node[c2], v2, n2, shard=r, n_shards=w)

    for key in ("weighted_unifrac", "unweighted_unifrac"):
        one, _ = _run_shards_in_threads(1, make2, key, computes=1)
        four, tms = _run_shards_in_threads(4, make2, key)
        assert np.array_equal(one, four), key
        assert all(t["exchange_bytes"] == 0 for t in tms)
    # Faith's PD with collectives installed: no vote, no exchange
    import threading

    from skbio_b200.distributed import ThreadCollectives

    coll = ThreadCollectives(2, timeout=60.0)
    pds, errs = [None, None], []

    def work(r):
        try:
            with make(r, 2) as prob:
                prob.set_collectives(r, 2, coll.for_rank(r))
                prob.compute("faith_pd")
                pds[r] = prob.sample_vector()
        except BaseException as e:
            errs.append(e)
            coll.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert np.array_equal(pds[0], pds[1])


def test_partial_upload_without_collectives_is_an_error():
    _skbio_b200()
    from skbio_b200 import _capi

    n, tips = 300, 500
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.05, 94)
    try:
        _capi.set_option("partial_upload", 1)
        with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n, shard=0, n_shards=2) as prob:
            with pytest.raises(_capi.EngineError, match="partial_upload"):
                prob.compute("weighted_unifrac")
        # a single shard is never partial
        with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
            prob.compute("weighted_unifrac")
    finally:
        _capi.set_option("partial_upload", 0)


def test_distributed_build_with_flagged_pairs_falls_back_per_rank():
    """Near-duplicate samples: the fixed-point guard flags their pairs, and recomputing a flagged pair
    needs the embedding of both samples, which a distributed build does not hold.  The shard that sees
    one redoes its own compute replicated (the others are not involved) and the result still meets
    the tolerance; the next compute votes for a replicated build everywhere."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 600, 1500
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.05, 93)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    big = 10 ** 7
    for a, b in ((5, 400), (130, 131), (300, 590)):   # b = a scaled, one count off: distance ~1e-7 of U
        dense[b] = dense[a] * big
        dense[b, np.flatnonzero(dense[a])[0]] += 1
    import scipy.sparse as sp

    csr = sp.csr_matrix(dense)
    indptr, cols, vals = csr.indptr.astype(np.int64), csr.indices.astype(np.int32), csr.data.astype(np.int64)
    exp = oracle_fast.unifrac_all(dense, tip_node, parent, length)
    for key in ("weighted_unifrac_normalized", "unweighted_unifrac"):
        def make(r, w):
            return _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n, shard=r, n_shards=w)

        base, tm = _run_shards_in_threads(1, make, key, computes=1)
        got, tms = _run_shards_in_threads(2, make, key, computes=3)
        assert_close_rel(got, exp[key], UNIFRAC_RTOL, key + " distributed build with near-duplicates")
        if tm[0]["tips_flagged"]:
            assert sum(t["tips_flagged"] for t in tms) == tm[0]["tips_flagged"]
            assert all(t["exchange_bytes"] == 0 for t in tms)   # third compute: everybody voted replicated


def test_devices_argument_on_two_real_gpus():
    """beta_diversity(..., devices=[0, 1]) on a box with two GPUs: one shard and one host thread per
    device, the per-block products exchanged by peer copies between the two devices, identical to the
    single-device result.  Skipped on a single-GPU box (devices=[0, 0] covers the logic there)."""
    sb = _skbio_b200()
    from skbio_b200 import _capi, synthetic

    if _capi.load().b2d_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    n, tips = 900, 2400
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.03, 97)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    tree = sb.FlatTree(parent, length, names)
    taxa = [f"T{i}" for i in range(tips)]
    for metric, kw in (("weighted_unifrac", {"normalized": True}), ("unweighted_unifrac", {}), ("braycurtis", {})):
        if "unifrac" in metric:
            kw = dict(kw, tree=tree, taxa=taxa)
        one = sb.beta_diversity(metric, dense, **kw).condensed_form()
        two = sb.beta_diversity(metric, dense, devices=[0, 1], **kw).condensed_form()
        assert np.array_equal(one, two, equal_nan=True), metric


def test_pcoa_centering_and_square_rows_on_the_resident_result():
    """Gower centering (pcoa's first step, /root/reference/skbio/stats/ordination/_cutils.pyx:19-123)
    and square-row gathering on the device-resident condensed result, against NumPy on the fetched
    vector (and against scikit-bio's own center_distance_matrix where the reference build is here)."""
    sb = _skbio_b200()
    from skbio_b200 import _capi
    from skbio_b200._dm import squareform

    n, tips = 333, 900
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.05, 33)
    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n) as prob:
        prob.compute("weighted_unifrac_normalized")
        vec = prob.fetch(np.empty(n * (n - 1) // 2))
        f, rm, gm = prob.center()
        rows = prob.fetch_square_rows(100, 229)
        text = __import__("io").StringIO()
        sb.write_lsmat(prob, text, ids=[f"s{i}" for i in range(n)], block_rows=100)
    sq = squareform(vec, n)
    assert np.array_equal(rows, sq[100:229])
    e = -0.5 * sq * sq
    erm = e.mean(axis=1)
    ef = e + ((e.mean() - erm)[:, None] - erm[None, :])
    assert np.allclose(rm, erm, rtol=1e-13, atol=0)
    assert abs(gm - e.mean()) <= 1e-13 * abs(e.mean())
    assert np.allclose(f, ef, rtol=0, atol=1e-13 * np.abs(ef).max())
    f2 = sb.center_distance_matrix(vec)
    assert np.array_equal(f2, f)
    lines = text.getvalue().split("\n")
    assert lines[0].split("\t")[1:] == [f"s{i}" for i in range(n)]
    assert lines[5].split("\t")[0] == "s4" and [float(x) for x in lines[5].split("\t")[1:]] == list(sq[4])
    from oracle import ref_skbio

    if ref_skbio.available():
        ref_skbio.load()
        from skbio.stats.ordination._principal_coordinate_analysis import center_distance_matrix

        assert np.allclose(f, center_distance_matrix(sq), rtol=0, atol=1e-13 * np.abs(ef).max())
    # a sharded problem has no whole result on one device
    with _capi.Problem.tree(parent, length, indptr, tip_node[cols], vals, n, shard=0, n_shards=2) as prob:
        prob.compute("weighted_unifrac")
        with pytest.raises(_capi.EngineError, match="whole triangle"):
            prob.center()


def test_sparse_table_50k_by_50k_through_the_public_api():
    """BASELINE config 3 at full size through beta_diversity() with a scipy.sparse table (stays
    sparse all the way to the device): Bray-Curtis and Jaccard, bit-exact on the pairs among 96
    samples spread over the table (SciPy on the densified rows), condensed result by default."""
    sb = _skbio_b200()
    import scipy.sparse as sp
    from scipy.spatial.distance import pdist

    from skbio_b200 import synthetic

    n = f = 50000
    indptr, cols, vals = synthetic.poisson_csr(n, f, 0.01, seed=0)
    table = sp.csr_matrix((vals, cols, indptr), shape=(n, f))
    pick = np.unique(np.concatenate([np.arange(0, 40), np.arange(24990, 25020), np.arange(n - 26, n)]))
    dense = np.asarray(table[pick].todense())
    for metric in ("braycurtis", "jaccard"):
        dm = sb.beta_diversity(metric, table)
        vec = dm.condensed_form()
        assert vec.shape == (n * (n - 1) // 2,)
        i, j = np.triu_indices(len(pick), k=1)
        gi, gj = pick[i], pick[j]
        got = vec[gi * n - ((gi + 2) * (gi + 1)) // 2 + gj]
        with np.errstate(all="ignore"):
            exp = pdist(dense > 0, "jaccard") if metric == "jaccard" else pdist(dense.astype(np.float64), "braycurtis")
        assert np.array_equal(got, exp, equal_nan=True), metric
        del dm, vec


def test_pipelined_intersection_kernel_on_feature_tables():
    """Bray-Curtis / Jaccard on sparse integer tables: the pipelined intersection kernel (32-bit sums)
    == the first-generation k_pair_sparse == SciPy, bit for bit; 5 000 features span two feature chunks,
    the dense-ish case cuts buckets into several batches."""
    _skbio_b200()
    from scipy.spatial.distance import pdist

    from skbio_b200 import _capi, synthetic

    for n, f, lam in ((400, 5000, 0.01), (300, 9000, 0.03)):
        indptr, cols, vals = synthetic.poisson_csr(n, f, lam, seed=7)
        dense = synthetic.csr_to_dense(indptr, cols, vals, f)
        with np.errstate(all="ignore"):
            exp = {"braycurtis": pdist(dense.astype(np.float64), "braycurtis"), "jaccard": pdist(dense > 0, "jaccard")}
        total = n * (n - 1) // 2
        try:
            _capi.set_option("sparse_features", 1)
            for kern, grp in ((0, 4), (1, 1), (1, 2), (1, 4), (1, 8)):
                _capi.set_option("isect_kernel", kern)
                _capi.set_option("isect_group", grp)
                for shards in (1, 2):
                    for key in ("braycurtis", "jaccard"):
                        out = np.full(total, -1.0)
                        for sh in range(shards):
                            with _capi.Problem.features(indptr, cols, vals.astype(np.float64), n, f, shard=sh,
                                                        n_shards=shards) as prob:
                                prob.compute(key)
                                assert prob.timings()["nodes_resident"] == 0   # the intersection path ran
                                prob.fetch(out)
                        assert np.array_equal(out, exp[key], equal_nan=True), (n, key, kern, grp, shards)
        finally:
            _capi.set_option("sparse_features", -1)
            _capi.set_option("isect_kernel", -1)
            _capi.set_option("isect_group", 0)


@pytest.mark.parametrize("cat", [False, True])
def test_unweighted_dense_rows_on_the_tensor_cores(cat):
    """Option udense_tc: the packed dense rows of unweighted UniFrac as an exact int8 contraction on
    tcgen05 (b2d_tc.cuh) instead of the look-up-table kernel - same distances (1e-13 vs the LUT path,
    1e-12 vs the oracle), for several sparse-row thresholds (few / many dense rows), shards, an exact
    copy (distance exactly 0) and a near-duplicate (flagged and recomputed)."""
    _skbio_b200()
    from oracle import oracle_fast
    from skbio_b200 import _capi, synthetic

    n, tips = 333, 4000
    parent, length, names, tip_node, indptr, cols, vals = _random_problem(n, tips, 0.015, 171, caterpillar=cat)
    rows = [cols[indptr[s]:indptr[s + 1]] for s in range(n)]
    rows[1] = rows[0].copy()
    cols2 = np.concatenate(rows)
    indptr2 = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    dense = np.zeros((n, tips), dtype=np.int64)
    for s in range(n):
        dense[s, rows[s]] = 1
    exp = oracle_fast.unifrac_all(dense, tip_node, parent, length)["unweighted_unifrac"]
    total = n * (n - 1) // 2
    res = {}
    try:
        for tc in (0, 1):
            _capi.set_option("udense_tc", tc)
            for k in (0, 1, 6, 200):
                _capi.set_option("sparse_rows_max_tips", k)
                for shards in (1, 2):
                    out = np.full(total, -1.0)
                    for sh in range(shards):
                        with _capi.Problem.tree(parent, length, indptr2, tip_node[cols2], None, n, shard=sh,
                                                n_shards=shards) as prob:
                            prob.compute("unweighted_unifrac")
                            assert prob.timings()["tips_ms"] > 0
                            prob.fetch(out)
                    res[(tc, k, shards)] = out
    finally:
        _capi.set_option("udense_tc", 0)
        _capi.set_option("sparse_rows_max_tips", 0)
    for key, v in res.items():
        assert v[0] == 0.0, key                                   # samples 0 and 1 are identical
        assert_close_rel(v, exp, UNIFRAC_RTOL, f"tensor-core dense rows {key} vs oracle")
        assert_close_rel(v, res[(0, key[1], 1)], 1e-13, f"tensor-core dense rows {key} vs LUT")
