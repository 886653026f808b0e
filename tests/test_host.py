"""Host-side logic (no GPU): tree flattening in reference id order, validation errors,
table ingestion, CSR conversion, the C-ABI library's exports and the engine's node order."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, case_arrays, golden_cases

import skbio_b200
from skbio_b200 import _capi, _table, _tree, synthetic
from skbio_b200._driver import beta_diversity


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_flatten_matches_reference_to_array(case):
    """names / lengths / parent in exactly the order TreeNode.to_array produced."""
    _, parent, length, names = case_arrays(case)
    tree = skbio_b200.read_newick(case["newick"])
    p, l, nm = skbio_b200.flatten_tree(tree)
    assert np.array_equal(p, parent)
    assert np.array_equal(l, length)
    assert list(nm) == list(names)
    assert all(p[k] > k for k in range(len(p) - 1)) and p[-1] == -1
    assert tree.id == len(p) - 1  # ids are written on the nodes like the reference does


def test_newick_round_trip_and_quotes():
    t = skbio_b200.read_newick("(('a b':1,'it''s':2)x:0.5,c_d:3[comment])r;")
    p, l, nm = skbio_b200.flatten_tree(t)
    assert list(nm) == ["a b", "it's", "x", "c d", "r"]
    assert list(l) == [1.0, 2.0, 0.5, 3.0, 0.0]
    parent, length, names, _ = synthetic.random_join_tree(50, seed=4)
    p2, l2, n2 = skbio_b200.flatten_tree(skbio_b200.read_newick(synthetic.to_newick(parent, length, names)))
    assert np.array_equal(p2, parent) and np.array_equal(l2, length) and list(n2) == list(names)
    p3, l3, n3 = skbio_b200.flatten_tree(synthetic.to_nodes(parent, length, names))
    assert np.array_equal(p3, parent) and np.array_equal(l3, length)


T1 = "(((((OTU1:0.5,OTU2:0.5):0.5,OTU3:1.0):1.0):0.0,(OTU4:0.75,OTU5:0.75):1.25):0.0)root;"
OIDS = ["OTU%d" % i for i in range(1, 6)]
B1 = [[1, 3, 0, 1, 0], [0, 2, 0, 4, 4]]


def test_validation_errors_match_reference_messages():
    """/root/reference/skbio/tree/_utils.py:61-90 and
    /root/reference/skbio/diversity/tests/test_driver.py:454-561."""
    t = skbio_b200.read_newick(T1)
    for metric in ("unweighted_unifrac", "weighted_unifrac"):
        with pytest.raises(ValueError, match="`taxa` is required"):
            beta_diversity(metric, B1, tree=t)
        with pytest.raises(ValueError, match="`tree` is required"):
            beta_diversity(metric, B1, taxa=OIDS)
        with pytest.raises(ValueError, match="All taxa must be unique"):
            beta_diversity(metric, B1, tree=t, taxa=["OTU1", "OTU1", "OTU3", "OTU4", "OTU5"])
        with pytest.raises(skbio_b200.MissingNodeError, match="1 taxa are not present as tip names"):
            beta_diversity(metric, B1, tree=t, taxa=["OTU1", "OTU2", "OTU3", "OTU4", "nope"])
        unrooted = skbio_b200.read_newick("((OTU1:0.1,OTU2:0.2):0.3,OTU3:0.5,(OTU4:0.7,OTU5:0.1):1.1);")
        with pytest.raises(ValueError, match="The tree must be rooted"):
            beta_diversity(metric, B1, tree=unrooted, taxa=OIDS)
        nolen = skbio_b200.read_newick("((OTU1:0.1,OTU2:0.2):0.3,((OTU3:0.5,OTU4:0.7),OTU5:1.1):1)root;")
        with pytest.raises(ValueError, match="must have a branch length"):
            beta_diversity(metric, B1, tree=nolen, taxa=OIDS)
        dup = skbio_b200.read_newick("((OTU1:0.1,OTU2:0.2):0.3,((OTU3:0.5,OTU4:0.7):1,(OTU5:1.1,OTU2:1):1):1)root;")
        with pytest.raises(skbio_b200.DuplicateNodeError, match="All tip names in the tree must be unique"):
            beta_diversity(metric, B1, tree=dup, taxa=OIDS)
        with pytest.raises(TypeError, match="unexpected keyword argument"):
            beta_diversity(metric, B1, tree=t, taxa=OIDS, not_a_known_parameter=1)
    with pytest.raises(ValueError, match="Counts cannot contain negative values"):
        beta_diversity("braycurtis", [[1, -2], [3, 4]])
    with pytest.raises(ValueError, match="Counts must be integers or floating-point numbers"):
        beta_diversity("braycurtis", np.array([["a", "b"], ["c", "d"]]))
    with pytest.raises(ValueError, match="is not an available beta diversity metric name"):
        beta_diversity("not-a-metric", [[1, 2], [3, 4]])
    with pytest.raises(ValueError, match="Input table has 2 samples whereas 3 sample IDs"):
        beta_diversity("braycurtis", [[1, 2], [3, 4]], ids=list("abc"))
    with pytest.raises(ValueError, match="Input table has 5 features whereas 2 feature IDs"):
        beta_diversity("unweighted_unifrac", B1, tree=t, taxa=["OTU1", "OTU2"])


def test_empty_input_gives_zero_matrix():
    dm = beta_diversity("braycurtis", np.zeros((3, 0)), ids=list("abc"))
    assert dm.shape == (3, 3) and tuple(dm.ids) == ("a", "b", "c")
    assert not np.asarray(dm.data).any()


def test_no_gpu_means_loud_failure_not_fallback():
    if skbio_b200.available():
        pytest.skip("a GPU is present")
    with pytest.raises(skbio_b200.EngineError):
        beta_diversity("braycurtis", [[1, 2], [3, 4]])
    with pytest.raises(skbio_b200.EngineError):
        beta_diversity("unweighted_unifrac", B1, tree=skbio_b200.read_newick(T1), taxa=OIDS)


def test_other_metrics_take_the_reference_route():
    from scipy.spatial.distance import pdist

    x = np.array([[1.0, 5.0], [2.0, 3.0], [0.0, 1.0]])
    dm = beta_diversity("euclidean", x, ids=list("abc"))
    assert np.allclose(dm.condensed_form(), pdist(x, "euclidean"))
    dm = beta_diversity("manhattan", x)
    assert np.allclose(dm.condensed_form(), pdist(x, "cityblock"))
    assert tuple(dm.ids) == ("0", "1", "2")
    dm = beta_diversity(lambda u, v: float(np.abs(u - v).max()), x)
    assert np.allclose(dm.condensed_form(), pdist(x, "chebyshev"))


def test_ingest_table_variants():
    import pandas as pd
    import scipy.sparse as sp

    df = pd.DataFrame([[1, 0, 2], [0, 3, 0]], index=["s1", "s2"], columns=["a", "b", "c"])
    data, sids, fids = _table.ingest_table(df)
    assert list(sids) == ["s1", "s2"] and list(fids) == ["a", "b", "c"] and data.shape == (2, 3)
    data, sids, fids = _table.ingest_table([1, 2, 3])
    assert data.shape == (1, 3) and sids is None
    with pytest.raises(TypeError, match="is not a supported table format"):
        _table.ingest_table(5)
    m = sp.csr_matrix(np.array([[0, 2.5, 0], [1.0, 0, 0.2]]))
    data, _, _ = _table.ingest_table(m)
    for mode, exp_vals in [("presence", None), ("int64", [2, 1]), ("float64", [2.5, 1.0, 0.2])]:
        ip, cols, vals = _table.to_csr(data, mode)
        ip2, cols2, vals2 = _table.to_csr(m.toarray(), mode)
        assert np.array_equal(ip, ip2) and np.array_equal(cols, cols2)
        if exp_vals is None:
            assert vals is None and vals2 is None
        else:
            assert list(vals) == exp_vals and list(vals2) == exp_vals


def test_taxa_lookup_last_node_wins_and_unobserved_skipped():
    names = np.array(["a", "b", "x", "a", None], dtype=object)
    out = _tree.map_taxa_to_nodes(["a", "b", "zz"], [0, 1], names)
    assert list(out) == [3, 1, -1]
    with pytest.raises(KeyError):
        _tree.map_taxa_to_nodes(["a", "zz"], [0, 1], names)


def test_library_exports_every_declared_symbol():
    """Every function declared in include/skbio_b200.h is exported by libskbb200.so."""
    header = open(os.path.join(ROOT, "include", "skbio_b200.h")).read()
    declared = set(re.findall(r"\b(b2d_[a-z_0-9]+)\s*\(", header))
    declared -= {"b2d_problem"}
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_capi.library_path())
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert _capi.load().b2d_api_version() == 1


def _emulate_propagation(parent, length, own):
    """Python model of the per-lane stack walk the propagation kernels execute."""
    t = _capi.debug_tree_order(parent, length)
    q_of, flags = t["q_of_id"], t["flags"]
    own_q = np.zeros(len(flags), dtype=np.int64)
    own_q[q_of] = own
    stack, out, maxd = [], np.zeros(len(flags), dtype=np.int64), 0
    for q in range(len(flags)):
        f = flags[q]
        if f & 4:
            continue
        v = own_q[q]
        if f & 1:
            v += stack.pop()
        out[q] = v
        if f & 2:
            stack[-1] += v
        else:
            stack.append(v)
        maxd = max(maxd, len(stack))
    assert maxd == t["max_depth"]
    return out[q_of], t


@pytest.mark.parametrize("gen,tips", [("random_join_tree", 257), ("caterpillar_tree", 2000),
                                      ("random_join_tree", 1), ("random_join_tree", 2)])
def test_engine_node_order_propagates_like_traverse_reduce(gen, tips):
    parent, length, names, tip_node = getattr(synthetic, gen)(tips, seed=5)
    rng = np.random.default_rng(0)
    own = np.zeros(len(parent), dtype=np.int64)
    own[tip_node] = rng.poisson(1.0, size=tips)
    got, t = _emulate_propagation(parent, length, own)
    exp = own.copy()
    for k in range(len(parent)):
        if parent[k] >= 0:
            exp[parent[k]] += exp[k]
    assert np.array_equal(got, exp)
    assert t["max_depth"] <= int(np.log2(len(parent))) + 2  # heavy-child-first bound
    assert sorted(t["q_of_id"]) == list(range(len(parent)))
    assert np.array_equal(t["len_q"][t["q_of_id"]], length)


@pytest.mark.parametrize("gen,tips", [("random_join_tree", 300), ("caterpillar_tree", 3000),
                                      ("random_join_tree", 1)])
def test_subtree_ranges_and_double_double_root_distances(gen, tips):
    """Arrays behind the closed-form part of zero-block skipping: the subtree of position q is
    the position range [q - size + 1, q], and hi + lo is the root distance to ~2^-100."""
    from fractions import Fraction

    parent, length, names, tip_node = getattr(synthetic, gen)(tips, seed=11)
    rng = np.random.default_rng(3)
    length = length * rng.uniform(1e-6, 1e3, size=len(length))  # lengths of very different size
    t = _capi.debug_tree_order(parent, length)
    pth = _capi.debug_tree_paths(parent, length)
    q_of = t["q_of_id"]
    m = len(parent)
    exact = [None] * m
    size = np.ones(m, dtype=np.int64)
    for k in range(m - 1, -1, -1):   # parents before children
        exact[k] = Fraction(float(length[k])) + (exact[parent[k]] if parent[k] >= 0 else 0)
    for k in range(m):
        if parent[k] >= 0:
            size[parent[k]] += size[k]
    lo_pos = np.full(m, m, dtype=np.int64)   # smallest position inside each node's subtree
    for k in range(m):
        lo_pos[k] = min(lo_pos[k], q_of[k])
        if parent[k] >= 0:
            lo_pos[parent[k]] = min(lo_pos[parent[k]], lo_pos[k])
    for k in range(m):
        q = q_of[k]
        assert pth["size_q"][q] == size[k]
        assert lo_pos[k] == q - size[k] + 1          # contiguous range ending at q
        assert pth["parent_q"][q] == (q_of[parent[k]] if parent[k] >= 0 else -1)
        dd = Fraction(float(pth["rd_hi"][q])) + Fraction(float(pth["rd_lo"][q]))
        assert abs(dd - exact[k]) <= exact[k] * Fraction(1, 2 ** 95)
        assert abs(pth["rd_lo"][q]) <= abs(pth["rd_hi"][q]) * 2.0 ** -52
    assert np.all(pth["parent_q"][m:] == -1)
    # tips below every node, and the root distance without the tips' own branches
    ntips = np.zeros(m, dtype=np.int64)
    has_child = np.zeros(m, dtype=bool)
    has_child[parent[parent >= 0]] = True
    ntips[~has_child] = 1
    for k in range(m):
        if parent[k] >= 0:
            ntips[parent[k]] += ntips[k]
    for k in range(m):
        q = q_of[k]
        assert pth["tips_q"][q] == ntips[k]
        if has_child[k] or parent[k] < 0:
            ref = (pth["rd_hi"][q], pth["rd_lo"][q]) if has_child[k] else (0.0, 0.0)
        else:
            pq = q_of[parent[k]]
            ref = (pth["rd_hi"][pq], pth["rd_lo"][pq])
        assert (pth["rd_hi_int"][q], pth["rd_lo_int"][q]) == ref
    assert np.all(pth["tips_q"][m:] == 0)


def test_engine_rejects_bad_trees():
    with pytest.raises(_capi.EngineError, match="parent"):
        _capi.debug_tree_order(np.array([0, -1]), np.zeros(2))
    with pytest.raises(_capi.EngineError, match="root"):
        _capi.debug_tree_order(np.array([-1, -1]), np.zeros(2))


def test_distance_matrix_stand_in():
    from skbio_b200._dm import SimpleDistanceMatrix, DistanceMatrixError

    dm = SimpleDistanceMatrix(np.array([1.0, 2.0, 3.0]), ["a", "b", "c"])
    assert dm.shape == (3, 3) and dm["a", "c"] == 2.0 and dm["c", "b"] == 3.0
    assert np.array_equal(dm.condensed_form(), [1.0, 2.0, 3.0])
    assert tuple(SimpleDistanceMatrix(np.zeros(0)).ids) == ("0",)
    with pytest.raises(DistanceMatrixError, match="IDs must be unique"):
        SimpleDistanceMatrix(np.zeros(3), ["a", "a", "b"])
    with pytest.raises(DistanceMatrixError, match="must match"):
        SimpleDistanceMatrix(np.zeros(3), ["a", "b"])


def test_synthetic_generators():
    parent, length, names, tip_node = synthetic.random_join_tree(100, seed=1)
    assert len(parent) == 199 and (parent[:-1] > np.arange(198)).all()
    assert length[-1] == 0.0 and (length[:-1] >= 0.01).all()
    indptr, cols, vals = synthetic.poisson_csr(50, 1000, 0.01, seed=0)
    assert indptr[-1] == len(cols) == len(vals) and (vals > 0).all() and cols.max() < 1000
    for s in range(50):
        r = cols[indptr[s]:indptr[s + 1]]
        assert (np.diff(r) > 0).all()
    p2, _, _, _ = synthetic.caterpillar_tree(300, seed=2)
    depth = np.zeros(len(p2), dtype=int)
    for k in range(len(p2) - 2, -1, -1):
        depth[k] = depth[p2[k]] + 1
    assert depth.max() > 60


def test_partial_beta_diversity_host_contract():
    """Errors and the callable route of partial_beta_diversity
    (/root/reference/skbio/diversity/tests/test_driver.py:814-951)."""
    from skbio_b200 import partial_beta_diversity

    t = skbio_b200.read_newick(T1)
    ids = ["A", "B"]
    with pytest.raises(ValueError, match="`id_pairs` are not a subset of `ids`"):
        partial_beta_diversity("unweighted_unifrac", B1, ids, [("A", "Z")], tree=t, taxa=OIDS)
    with pytest.raises(ValueError, match="A duplicate or a self-self pair was observed"):
        partial_beta_diversity("unweighted_unifrac", B1, ids, [("A", "B"), ("B", "A")], tree=t, taxa=OIDS)
    with pytest.raises(ValueError, match="A duplicate or a self-self pair was observed"):
        partial_beta_diversity("unweighted_unifrac", B1, ids, [("A", "A")], tree=t, taxa=OIDS)
    with pytest.raises(ValueError, match="only compatible with optimized unifrac methods"):
        partial_beta_diversity("braycurtis", B1, ids, [("A", "B")])
    dm = partial_beta_diversity(lambda u, v: float(np.abs(u - v).sum()), np.asarray(B1), ids, [("B", "A")])
    assert dm["A", "B"] == 9.0 and dm["B", "A"] == 9.0


def test_real_skbio_treenode_is_accepted_when_available():
    """Duck typing against the real classes (only where scikit-bio is importable, e.g. the build
    container with the out-of-tree reference build on PYTHONPATH): flatten_tree(TreeNode) equals
    TreeNode.to_array, validation raises the reference's own exception types."""
    skbio = pytest.importorskip("skbio")
    from io import StringIO

    case = [c for c in golden_cases() if c["name"] == "multifurcating_unary_rootlen"][0]
    t = skbio.TreeNode.read(StringIO(case["newick"]))
    p, l, nm = skbio_b200.flatten_tree(t)
    arr = t.to_array(nan_length_value=0.0)
    assert list(nm) == list(arr["name"]) and np.array_equal(l, arr["length"])
    assert np.array_equal(p, np.array(case["tree_parent"]))
    from skbio.tree import MissingNodeError

    with pytest.raises(skbio_b200.MissingNodeError):
        skbio_b200.validate_taxa_and_tree(["a", "nope"], t)
    # the reference's own exception types are used when scikit-bio was importable at import time
    # (in one pytest process the reference build may only have been put on sys.path later)
    if skbio_b200.MissingNodeError.__module__.startswith("skbio."):
        assert skbio_b200.MissingNodeError is MissingNodeError


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_native_newick_parser_matches_reference_to_array(case):
    """libskbb200's Newick -> flat arrays equals TreeNode.read + to_array of the reference."""
    _, parent, length, names = case_arrays(case)
    ft = skbio_b200.read_newick_flat(case["newick"])
    p, l, nm = skbio_b200.flatten_tree(ft)
    assert np.array_equal(p, parent) and np.array_equal(l, length) and list(nm) == list(names)


def test_native_newick_parser_details(qiime_tt):
    ft = skbio_b200.read_newick_flat("(('a b':1,'it''s':2)x:0.5,c_d:3[comment [nested]])r;")
    assert list(ft.names) == ["a b", "it's", "x", "c d", "r"]
    assert np.isnan(ft.length[-1]) and list(ft.length[:-1]) == [1.0, 2.0, 0.5, 3.0]
    ft = skbio_b200.read_newick_flat(" ( a : 1e-3 ,\n b:2 ) ; trailing junk")
    assert list(ft.names) == ["a", "b", None] and ft.length[0] == 1e-3
    py = skbio_b200.flatten_tree(skbio_b200.read_newick(qiime_tt["newick"]))
    nat = skbio_b200.flatten_tree(skbio_b200.read_newick_flat(qiime_tt["newick"]))
    assert all(np.array_equal(a, b) for a, b in zip(py[:2], nat[:2])) and list(py[2]) == list(nat[2])
    for bad in ["((a,b);", "(a,b", "(a:x,b);", "('a,b);", "(a b,c);"]:
        with pytest.raises(_capi.EngineError):
            skbio_b200.read_newick_flat(bad)
    # big random tree round trip
    parent, length, names, _ = synthetic.random_join_tree(3000, seed=9)
    ft = skbio_b200.read_newick_flat(synthetic.to_newick(parent, length, names))
    assert np.array_equal(ft.parent, parent) and list(ft.names) == list(names)
    assert np.array_equal(ft.length[:-1], length[:-1]) and np.isnan(ft.length[-1])


def test_flat_tree_validation_messages():
    t = skbio_b200.read_newick_flat(T1)
    skbio_b200.validate_taxa_and_tree(OIDS, t)
    with pytest.raises(skbio_b200.MissingNodeError, match="1 taxa are not present"):
        skbio_b200.validate_taxa_and_tree(OIDS + ["nope"], t)
    with pytest.raises(ValueError, match="The tree must be rooted"):
        skbio_b200.validate_taxa_and_tree(OIDS, skbio_b200.read_newick_flat(
            "((OTU1:0.1,OTU2:0.2):0.3,OTU3:0.5,(OTU4:0.7,OTU5:0.1):1.1);"))
    with pytest.raises(ValueError, match="must have a branch length"):
        skbio_b200.validate_taxa_and_tree(OIDS, skbio_b200.read_newick_flat(
            "((OTU1:0.1,OTU2:0.2):0.3,((OTU3:0.5,OTU4:0.7),OTU5:1.1):1)root;"))
    with pytest.raises(skbio_b200.DuplicateNodeError):
        skbio_b200.validate_taxa_and_tree(OIDS, skbio_b200.read_newick_flat(
            "((OTU1:0.1,OTU2:0.2):0.3,((OTU3:0.5,OTU4:0.7):1,(OTU5:1.1,OTU2:1):1):1)root;"))


def test_native_scan_of_stored_values_matches_numpy():
    """b2d_host_scan (the negativity check of _validate_counts_matrix and the 'every stored entry counts'
    test of the CSR conversion in one multi-threaded pass): minimum, non-zero and positive counts for the
    four dtypes, NaN included, small arrays through the NumPy branch."""
    from skbio_b200 import _capi, _table

    rng = np.random.default_rng(5)
    for dt in (np.int64, np.float64, np.int32, np.float32):
        for n in (0, 3, 70001, 1_200_003):
            if np.issubdtype(dt, np.integer):
                x = rng.integers(-2, 7, n).astype(dt)
            else:
                x = (rng.normal(size=n) * (rng.random(n) < 0.6)).astype(dt)
                if n > 10:
                    x[5] = np.nan
            mn, nz, pos = _capi.host_scan(x)
            assert nz == int(np.count_nonzero(x)) and pos == int(np.count_nonzero(x > 0))
            if n:
                with np.errstate(all="ignore"):
                    assert mn == float(np.nanmin(x))
    import scipy.sparse as sp

    bad = sp.csr_matrix(np.array([[0, 2, -1], [3, 0, 0]]))
    with pytest.raises(ValueError, match="negative"):
        _table.validate_counts_matrix(bad)
    big = sp.random(300, 2000, density=0.2, random_state=3, format="csr", dtype=np.float64)
    big.data[17] = 0.0      # an explicit zero must not become an entry
    indptr, cols, vals = _table.to_csr(big, "float64")
    assert len(vals) == big.nnz - 1 and indptr[-1] == len(vals) and np.all(vals != 0)
