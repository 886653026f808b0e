"""The unmodified reference (oracle/_ref, built by tools/build_ref.sh) against the oracle
restatements and the host half of the engine.  CPU only; skipped where the build is absent
(a fresh checkout outside the build container: oracle/_ref is git-ignored build output)."""
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_skbio  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_skbio.available(), reason="oracle/_ref not built (tools/build_ref.sh)")


def _problem(n, tips, lam, seed, caterpillar=False):
    from skbio_b200 import synthetic

    gen = synthetic.caterpillar_tree if caterpillar else synthetic.random_join_tree
    parent, length, names, tip_node = gen(tips, seed=seed + 1)
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=seed)
    dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
    return parent, length, names, tip_node, dense, synthetic.to_newick(parent, length, names)


@pytest.mark.parametrize("cat", [False, True])
def test_c_oracle_is_bit_equal_to_the_reference(cat):
    """oracle_c.c (the checker of every GPU parity test) == skbio.diversity.beta_diversity of the
    reference build, all five metrics, bit for bit, on a seeded random case with an empty sample
    and a duplicated one."""
    skbio = ref_skbio.load()
    from skbio.diversity import beta_diversity

    from oracle import oracle_fast

    n, tips = 60, 300
    parent, length, names, tip_node, dense, newick = _problem(n, tips, 0.05, 5, cat)
    dense[7] = 0
    dense[9] = dense[8]
    tree = skbio.TreeNode.read(io.StringIO(newick))
    taxa = [f"T{i}" for i in range(tips)]
    got = oracle_fast.unifrac_all(dense, tip_node, parent, length)
    with np.errstate(all="ignore"):
        exp = {
            "unweighted_unifrac": beta_diversity("unweighted_unifrac", dense, tree=tree, taxa=taxa),
            "weighted_unifrac": beta_diversity("weighted_unifrac", dense, tree=tree, taxa=taxa),
            "weighted_unifrac_normalized": beta_diversity("weighted_unifrac", dense, tree=tree, taxa=taxa,
                                                          normalized=True),
            "braycurtis": beta_diversity("braycurtis", dense),
            "jaccard": beta_diversity("jaccard", dense),
        }
    iu = np.triu_indices(n, k=1)
    for key, dm in exp.items():
        assert np.array_equal(got[key], np.asarray(dm.data)[iu], equal_nan=True), key


def test_host_arrays_follow_the_reference_for_duplicate_taxa():
    """Several observed columns mapped to one node (validate=False): the reference assigns whole
    columns in ascending order (_phylogenetic.pyx:208-218), so the LAST observed column decides every
    sample's count - a sample that is non-zero only in an earlier duplicate ends up with 0."""
    skbio = ref_skbio.load()
    from skbio.diversity._util import vectorize_counts_and_tree

    import skbio_b200
    from skbio_b200 import _driver

    newick = "((a:1,b:2)x:0.5,(c:1,d:3)y:0.25)r;"
    taxa = ["a", "b", "a", "c", "d", "c"]
    counts = np.array([[5, 1, 0, 2, 0, 0],     # 'a': later column is 0 -> 0; 'c': later column 0 -> 0
                       [0, 0, 7, 0, 1, 4],
                       [3, 2, 9, 1, 1, 0],
                       [0, 4, 0, 0, 0, 0]])
    tree = skbio.TreeNode.read(io.StringIO(newick))
    exp, _, _ = vectorize_counts_and_tree(counts, np.asarray(taxa), tree)
    exp = np.asarray(exp)
    flat = skbio_b200.read_newick_flat(newick)
    p, l, ip, ids, vals, n = _driver._unifrac_host_arrays("weighted_unifrac", counts, taxa, flat, False)
    m = len(p)
    emb = np.zeros((n, m), dtype=np.int64)
    emb[np.repeat(np.arange(n), np.diff(ip)), ids] = vals
    for k in range(m - 1):
        emb[:, p[k]] += emb[:, k]
    assert np.array_equal(emb, exp)


def test_flat_newick_reader_equals_treenode_to_array():
    skbio = ref_skbio.load()
    import skbio_b200

    parent, length, names, tip_node, dense, newick = _problem(4, 500, 0.1, 11)
    arr = skbio.TreeNode.read(io.StringIO(newick)).to_array(nan_length_value=0.0)
    p, l, nm = skbio_b200.flatten_tree(skbio_b200.read_newick_flat(newick))
    assert np.array_equal(l, arr["length"])
    assert list(nm) == list(arr["name"])


def test_streaming_lsmat_writer_produces_the_reference_file():
    """write_lsmat (row blocks from a condensed vector) == DistanceMatrix.write(format='lsmat') of
    the reference, byte for byte, including awkward values (0.1 + 0.2, 1e-5, 1/3)."""
    skbio = ref_skbio.load()
    import skbio_b200

    rng = np.random.default_rng(3)
    n = 37
    vec = rng.random(n * (n - 1) // 2)
    vec[:4] = [0.1 + 0.2, 1e-5, 1.0 / 3.0, 0.0]
    ids = [f"sample.{i}" for i in range(n)]
    dm = skbio.DistanceMatrix(vec, ids)
    ref_txt = io.StringIO()
    dm.write(ref_txt, format="lsmat")
    for block in (1, 5, 64):
        got = io.StringIO()
        assert skbio_b200.write_lsmat(vec, got, ids=ids, block_rows=block) == n
        assert got.getvalue() == ref_txt.getvalue()
    got = io.StringIO()
    skbio_b200.write_lsmat(dm, got)          # ids taken from the object
    assert got.getvalue() == ref_txt.getvalue()


def test_centering_formula_against_the_reference_numpy_path():
    """The host statement of what the device computes for pcoa centering (E, row means, F) against
    skbio.stats.ordination's own center_distance_matrix - the GPU test checks the kernels against
    the same function."""
    skbio = ref_skbio.load()
    from skbio.stats.ordination._principal_coordinate_analysis import center_distance_matrix

    rng = np.random.default_rng(4)
    n = 23
    vec = rng.random(n * (n - 1) // 2)
    sq = np.asarray(skbio.DistanceMatrix(vec).data)
    exp = center_distance_matrix(sq)
    e = -0.5 * sq * sq
    rm = e.mean(axis=1)
    f = e + ((e.mean() - rm)[:, None] - rm[None, :])
    assert np.allclose(f, exp, rtol=0, atol=1e-14)
