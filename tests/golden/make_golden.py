"""Generate golden fixtures by running the UNMODIFIED reference (scikit-bio) in the build
container.  The reference cannot travel to the GPU box, so its outputs are committed here.

Usage (build container only):

    # 1. build the reference out of tree
    mkdir -p /tmp/refbuild && cd /tmp/refbuild
    cp -r /root/reference/{skbio,setup.py,pyproject.toml,README.rst,LICENSE.txt} .
    DISABLE_OPENMP=1 python setup.py build_ext --inplace
    # 2. shim the missing third-party imports (h5py, biom, natsort, ...): the
    #    sitecustomize.py of SURVEY.md appendix A in /tmp/refshim
    # 3. run
    PYTHONPATH=/tmp/refshim:/tmp/refbuild:/root/repo python tests/golden/make_golden.py

Writes ``tests/golden/cases.json`` (inputs + reference outputs for every case) and
``tests/golden/qiime191tt.json`` (the QIIME 1.9.1 "tiny test" table/tree/expected matrices
from /root/reference/skbio/diversity/beta/tests/data/qiime-191-tt/).

Every output value in these files was produced by ``skbio.diversity.beta_diversity`` /
``TreeNode.to_array`` / ``vectorize_counts_and_tree`` of the reference, not by this repo.
"""

import json
import os
from io import StringIO

import numpy as np

import skbio
from skbio import TreeNode
from skbio.diversity import beta_diversity, alpha_diversity
from skbio.diversity._util import vectorize_counts_and_tree

from skbio_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
REFDATA = "/root/reference/skbio/diversity/beta/tests/data/qiime-191-tt"

METRICS = [
    ("unweighted_unifrac", {}),
    ("weighted_unifrac", {"normalized": False}),
    ("weighted_unifrac", {"normalized": True}),
    ("braycurtis", {}),
    ("jaccard", {}),
]


def run_case(name, newick, counts, taxa, ids=None, dtype="int"):
    counts = np.asarray(counts)
    tree = TreeNode.read(StringIO(newick))
    arr = tree.to_array(nan_length_value=0.0)
    m = len(arr["id"])
    parent = np.full(m, -1, dtype=np.int64)
    for node in tree.traverse(include_self=True):
        if node.parent is not None and node is not tree:
            parent[node.id] = node.parent.id
    out = {
        "name": name,
        "newick": newick,
        "counts": counts.tolist(),
        "counts_dtype": str(counts.dtype),
        "taxa": list(taxa),
        "ids": None if ids is None else list(ids),
        "tree_names": [None if x is None else str(x) for x in arr["name"]],
        "tree_length": [float(x) for x in arr["length"]],
        "tree_parent": parent.tolist(),
        "results": {},
    }
    cbn, _, _ = vectorize_counts_and_tree(counts, np.asarray(taxa), tree)
    out["counts_by_node"] = np.asarray(cbn).tolist()
    for metric, kw in METRICS:
        key = metric + ("_normalized" if kw.get("normalized") else "")
        kwargs = dict(kw)
        if "unifrac" in metric:
            kwargs.update(tree=tree, taxa=taxa)
        with np.errstate(all="ignore"):
            dm = beta_diversity(metric, counts, ids, **kwargs)
        out["results"][key] = [float(x) for x in dm.condensed_form()]
        out["result_ids"] = list(dm.ids)
    with np.errstate(all="ignore"):
        fp = alpha_diversity("faith_pd", counts.astype(int) if counts.dtype != bool else counts,
                             ids, tree=tree, taxa=taxa)
    out["results"]["faith_pd"] = [float(x) for x in fp.values]
    icounts = counts.astype(int) if counts.dtype != bool else counts
    for rooted in (True, False):
        for weight in (False, True, 0.5):
            with np.errstate(all="ignore"):
                pv = alpha_diversity("phydiv", icounts, ids, tree=tree, taxa=taxa, rooted=rooted,
                                     weight=weight, validate=False)
            out["results"][f"phydiv_r{int(rooted)}_w{float(weight)}"] = [float(x) for x in pv.values]
    return out


def random_case(name, n_samples, n_tips, lam, seed, n_taxa=None, floats=False,
                empty_rows=(), caterpillar=False, root_length=None):
    rng = np.random.default_rng(seed)
    gen = synthetic.caterpillar_tree if caterpillar else synthetic.random_join_tree
    parent, length, names, tip_node = gen(n_tips, seed=seed + 1)
    if root_length is not None:
        length = length.copy()
        length[-1] = root_length
    newick = synthetic.to_newick(parent, length, names, root_length_none=root_length is None)
    n_taxa = n_tips if n_taxa is None else n_taxa
    tips = rng.permutation(n_tips)[:n_taxa]
    taxa = [f"T{i}" for i in tips]
    counts = rng.poisson(lam, size=(n_samples, n_taxa))
    if floats:
        counts = counts * rng.uniform(0.2, 1.7, size=counts.shape)
        counts[counts < 0.3] = 0.0
    for r in empty_rows:
        counts[r, :] = 0
    ids = [f"S{i}" for i in range(n_samples)]
    return run_case(name, newick, counts, taxa, ids)


def main():
    cases = []
    # --- the reference's own unit-test inputs (test_unifrac.py:24-45) ---
    b1 = [[1, 3, 0, 1, 0], [0, 2, 0, 4, 4], [0, 0, 6, 2, 1],
          [0, 0, 1, 1, 1], [5, 3, 5, 0, 0], [0, 0, 0, 3, 5]]
    oids1 = ["OTU%d" % i for i in range(1, 6)]
    t1 = ("(((((OTU1:0.5,OTU2:0.5):0.5,OTU3:1.0):1.0):0.0,(OTU4:"
          "0.75,OTU5:0.75):1.25):0.0)root;")
    t1x = ("(((((OTU1:0.5,OTU2:0.5):0.5,OTU3:1.0):1.0):0.0,(OTU4:"
           "0.75,(OTU5:0.25,(OTU6:0.5,OTU7:0.5):0.5):0.5):1.25):0.0)root;")
    cases.append(run_case("b1_t1", t1, b1, oids1, list("ABCDEF")))
    cases.append(run_case("b1_t1_extra_tips", t1x, b1, oids1, list("ABCDEF")))
    t2 = "((OTU1:0.1, OTU2:0.2):0.3, (OTU3:0.5, OTU4:0.7):1.1)root;"
    oids2 = ["OTU%d" % i for i in range(1, 5)]
    cases.append(run_case("t2_root_not_observed",
                          t2, [[1, 1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 1], [0, 0, 1, 0],
                               [0, 0, 0, 0], [0, 0, 0, 0]], oids2))
    cases.append(run_case("minimal_two_tips", "(OTU1:0.25, OTU2:0.25)root;",
                          [[1, 0], [0, 0], [0, 1], [1, 1]], ["OTU1", "OTU2"]))
    # docstring example (_unifrac.py:125-140)
    doc_tree = ("(((((OTU1:0.5,OTU2:0.5):0.5,OTU3:1.0):1.0):0.0,"
                "(OTU4:0.75,(OTU5:0.5,((OTU6:0.33,OTU7:0.62):0.5"
                ",OTU8:0.5):0.5):0.5):1.25):0.0)root;")
    cases.append(run_case("docstring", doc_tree,
                          [[1, 0, 0, 4, 1, 2, 3, 0], [0, 1, 1, 6, 0, 1, 0, 0],
                           [5, 0, 0, 0, 0, 0, 0, 1]], ["OTU%d" % i for i in range(1, 9)]))
    # float counts: unweighted treats 0.5 as present, weighted truncates to 0
    cases.append(run_case("float_counts", t2,
                          np.array([[0.5, 1.5, 0.0, 2.0], [1.0, 0.0, 0.7, 0.0],
                                    [0.0, 0.0, 3.2, 1.0]]), oids2))
    # multifurcation, unary nodes, zero lengths, internal names, root with a length
    weird = "((a:1,b:2,c:0,d:0.5)x:0.25,((e:3)y:0,(f:0.125,g:1e-3)z:2)w:1.5)r:0.75;"
    rng = np.random.default_rng(7)
    cases.append(run_case("multifurcating_unary_rootlen", weird,
                          rng.poisson(0.8, size=(9, 7)), list("gfedcba")))
    # an internal node named like an observed taxon and numbered after the tip:
    # the name lookup picks the later (internal) node (_phylogenetic.pyx:195-199)
    collide = "((a:1,b:2)q:0.5,(c:1,(d:1,e:1)a:0.5)p:1)r;"
    cases.append(run_case("internal_name_collision", collide,
                          rng.poisson(1.0, size=(6, 5)) + np.eye(6, 5, dtype=int),
                          list("abcde")))
    # single sample -> 1x1
    cases.append(run_case("single_sample", t2, [[1, 2, 0, 1]], oids2))
    # random trees
    cases.append(random_case("rand_40x60", 40, 60, 0.3, seed=11, empty_rows=(5, 17)))
    cases.append(random_case("rand_subset_taxa", 33, 100, 0.5, seed=12, n_taxa=37))
    cases.append(random_case("rand_floats", 20, 50, 0.7, seed=13, floats=True, empty_rows=(0,)))
    cases.append(random_case("rand_caterpillar", 25, 200, 0.2, seed=14, caterpillar=True))
    cases.append(random_case("rand_rootlen", 17, 31, 1.5, seed=15, root_length=0.4))
    cases.append(random_case("rand_130x300", 130, 300, 0.05, seed=16, empty_rows=(1, 2, 129)))
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py",
                   "reference": "scikit-bio " + skbio.__version__,
                   "cases": cases}, f)

    # --- QIIME 1.9.1 tiny test ---
    import pandas as pd
    tab = pd.read_csv(os.path.join(REFDATA, "otu-table.tsv"), sep="\t", skiprows=1, index_col=0)
    tab = tab.drop(columns=[c for c in tab.columns if c == "not16S.1"])
    newick = open(os.path.join(REFDATA, "tree.nwk")).read().strip()
    q = {"newick": newick, "taxa": [str(x) for x in tab.index],
         "ids": [str(x) for x in tab.columns], "counts": tab.to_numpy().T.tolist(),
         "expected": {}, "reference_run": {}}
    tree = TreeNode.read(StringIO(newick))
    for fn, key, metric, kw in [
            ("unweighted_unifrac_dm.txt", "unweighted_unifrac", "unweighted_unifrac", {}),
            ("weighted_unifrac_dm.txt", "weighted_unifrac", "weighted_unifrac", {}),
            ("weighted_normalized_unifrac_dm.txt", "weighted_unifrac_normalized",
             "weighted_unifrac", {"normalized": True})]:
        exp = pd.read_csv(os.path.join(REFDATA, fn), sep="\t", index_col=0)
        exp = exp.loc[q["ids"], q["ids"]]
        q["expected"][key] = exp.to_numpy().tolist()
        dm = beta_diversity(metric, np.asarray(q["counts"]), q["ids"], tree=tree,
                            taxa=q["taxa"], **kw)
        q["reference_run"][key] = [float(x) for x in dm.condensed_form()]
    with open(os.path.join(HERE, "qiime191tt.json"), "w") as f:
        json.dump(q, f)
    # --- PERMANOVA (consumer of the distance matrix; reference: stats/distance/_permanova.py) ---
    from scipy.spatial.distance import pdist
    from skbio.stats.distance import DistanceMatrix, permanova
    pm = []
    for n, k, feats, perms, seed, effect in [(30, 3, 12, 99, 7, 2.0), (120, 4, 40, 199, 42, 0.05),
                                             (150, 2, 30, 99, 3, 0.0), (40, 5, 8, 0, 1, 1.0),
                                             (64, 3, 10, 499, 11, 0.0)]:
        rng = np.random.default_rng(seed + 1000)
        grouping = rng.integers(0, k, size=n)
        x = rng.poisson(3.0 + effect * grouping[:, None] * rng.random(feats)[None, :], size=(n, feats))
        vec = pdist(x, "braycurtis")
        dm = DistanceMatrix(vec, [f"s{i}" for i in range(n)])
        labels = [f"g{g}" for g in grouping]
        res = permanova(dm, labels, permutations=perms, seed=seed)
        pm.append({"n": n, "condensed": [float(v) for v in vec], "grouping": labels,
                   "permutations": perms, "seed": seed,
                   "stat": float(res["test statistic"]),
                   "p_value": None if perms == 0 else float(res["p-value"]),
                   "num_groups": int(res["number of groups"])})
    with open(os.path.join(HERE, "permanova.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": pm}, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
