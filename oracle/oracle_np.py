"""CPU oracle (NumPy) for the all-pairs beta-diversity path.  TEST INFRASTRUCTURE ONLY.

This file restates, step by step, what scikit-bio's CPU implementation computes for
``beta_diversity`` with ``unweighted_unifrac``, ``weighted_unifrac`` (normalized or not),
``braycurtis`` and ``jaccard``.  It is the checker the CUDA path is compared with.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it;
the product package (``scikit-bio_b200/``) never does.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here against
(a) the known-answer values in the reference's own tests (QIIME 1.9.1 / PyCogent / hand
values, /root/reference/skbio/diversity/beta/tests/test_unifrac.py:101-170, 319-687 and
/root/reference/skbio/diversity/tests/test_driver.py:625-713), (b) the
``qiime-191-tt`` distance matrices shipped in
/root/reference/skbio/diversity/beta/tests/data/qiime-191-tt/, and (c) outputs of the
reference itself, generated in the build container by ``tests/golden/make_golden.py`` and
committed as ``tests/golden/*.json``.

Third-party arithmetic: Bray-Curtis and Jaccard (and the pair enumeration order) are computed
by SciPy's ``pdist`` in the reference (``scipy>=1.9.0``, unpinned in
/root/reference/pyproject.toml:42; 1.18.1 in this image), called at
/root/reference/skbio/diversity/_driver.py:349-354.  SciPy's published definitions
(``scipy/spatial/distance.py`` ``braycurtis`` / ``jaccard``) are restated below and also
checked against the installed SciPy.

Inputs are flat arrays (the tree as ``parent``/``length``/``names`` in the reference's node
id order, see /root/reference/skbio/tree/_tree.py:5337-5359) so that the oracle is
independent of the product's tree flattener; ``tree_arrays_from_nested`` builds them from a
nested-tuple tree with its own, separate implementation of the reference's id assignment.
"""

from __future__ import annotations

from functools import partial

import numpy as np


# --------------------------------------------------------------------------------------
# tree indexing (restates TreeNode.assign_ids / index_tree / to_array)
# --------------------------------------------------------------------------------------
def tree_arrays_from_nested(nested):
    """Flat arrays for a tree given as nested tuples ``(name, length, [children...])``.

    Follows ``assign_ids`` (/root/reference/skbio/tree/_tree.py:5337-5359): walk the tree in
    postorder; whenever a node is reached, its children get the next consecutive ids; the
    root gets the last id.  ``length`` ``None`` becomes 0.0 as ``to_array(nan_length_value=
    0.0)`` does (_tree.py:6326-6328).  Returns ``(parent, length, names, child_index)`` with
    ``child_index`` rows ``(node, first_child, last_child)`` in postorder as ``index_tree``
    builds them (_tree.py:5361-5400).
    """
    # recursive on purpose: small trees only, and structurally different from the product
    ids = {}
    counter = [0]
    order = []  # postorder list of nested nodes

    def post(node):
        for c in node[2]:
            post(c)
        order.append(node)

    post(nested)
    for node in order:
        for c in node[2]:
            ids[id(c)] = counter[0]
            counter[0] += 1
    ids[id(nested)] = counter[0]
    m = counter[0] + 1
    parent = np.full(m, -1, dtype=np.int64)
    length = np.zeros(m, dtype=np.float64)
    names = np.empty(m, dtype=object)
    child_index = []
    for node in order:
        k = ids[id(node)]
        names[k] = node[0]
        length[k] = 0.0 if node[1] is None else node[1]
        if node[2]:
            for c in node[2]:
                parent[ids[id(c)]] = k
    for node in order:  # index_tree: rows appended when the *parent* is visited
        for c in node[2]:
            if c[2]:
                child_index.append((ids[id(c)], ids[id(c[2][0])], ids[id(c[2][-1])]))
    if nested[2]:
        child_index.append((ids[id(nested)], ids[id(nested[2][0])], ids[id(nested[2][-1])]))
    return parent, length, names, np.atleast_2d(np.asarray(child_index, dtype=np.int64))


def child_index_from_parent(parent):
    """``child_index`` rows from a parent array (siblings have consecutive ids)."""
    parent = np.asarray(parent)
    m = len(parent)
    first = np.full(m, -1, dtype=np.int64)
    last = np.full(m, -1, dtype=np.int64)
    for k in range(m):
        p = parent[k]
        if p >= 0:
            if first[p] < 0:
                first[p] = k
            last[p] = k
    rows = [(k, first[k], last[k]) for k in range(m) if first[k] >= 0]
    # the reference lists them in postorder of the parents; any order in which children
    # rows precede their parents' rows gives the same sums; ascending node id is one.
    return np.asarray(rows, dtype=np.int64).reshape(-1, 3)


# --------------------------------------------------------------------------------------
# _nodes_by_counts + _traverse_reduce  (/root/reference/skbio/diversity/_phylogenetic.pyx)
# --------------------------------------------------------------------------------------
def nodes_by_counts(counts, taxa, names, child_index):
    """Dense ``(n_nodes, n_samples)`` int64 array of per-node descendant counts.

    Restates ``_nodes_by_counts`` (_phylogenetic.pyx:148-222): cast to int64 (:186,
    truncating floats), observed taxa are the columns with a non-zero sum (:190), each
    observed taxon is mapped to the *last* node carrying its name (:195-199), tip rows are
    filled (:208-218), then ``_traverse_reduce`` (:73-143) adds children rows into their
    parent row following ``child_index``.
    """
    counts = np.atleast_2d(counts).astype(np.int64, copy=False)
    taxa = np.asarray(taxa, dtype=object)
    observed_indices = counts.sum(0).nonzero()[0]
    observed_ids = taxa[observed_indices]
    observed_set = set(observed_ids)
    node_lookup = {}
    for i in range(len(names)):
        if names[i] in observed_set:
            node_lookup[names[i]] = i
    a = np.zeros((len(names), counts.shape[0]), dtype=np.int64)
    for i, col in enumerate(observed_indices):
        a[node_lookup[observed_ids[i]], :] = counts[:, col]
    for node, start, end in child_index:
        if end >= start >= 0:
            a[node, :] += a[start : end + 1, :].sum(axis=0)
    return a


def tip_indices_from_parent(parent):
    parent = np.asarray(parent)
    has_child = np.zeros(len(parent), dtype=bool)
    has_child[parent[parent >= 0]] = True
    return np.nonzero(~has_child)[0]


def tip_distances(length, parent, tip_indices):
    """``_tip_distances`` (_phylogenetic.pyx:24-68): ``d[i] += d[parent[i]]`` in preorder,
    then everything that is not a tip is zeroed.  With reference ids ``parent[i] > i``, so
    descending id order visits parents before children, as preorder does."""
    d = np.array(length, dtype=np.float64, copy=True)
    for i in range(len(d) - 1, -1, -1):
        p = parent[i]
        if p >= 0:
            d[i] += d[p]
    mask = np.zeros(len(d), dtype=np.float64)
    mask[tip_indices] = 1.0
    return d * mask


# --------------------------------------------------------------------------------------
# pair functions (/root/reference/skbio/diversity/beta/_unifrac.py:340-471, 594-619)
# --------------------------------------------------------------------------------------
def unweighted_unifrac_pair(u_node_counts, v_node_counts, branch_lengths):
    unique_nodes = np.logical_xor(u_node_counts, v_node_counts)
    observed_nodes = np.logical_or(u_node_counts, v_node_counts)
    unique_branch_length = (branch_lengths * unique_nodes).sum()
    observed_branch_length = (branch_lengths * observed_nodes).sum()
    if observed_branch_length == 0.0:
        return 0.0
    return unique_branch_length / observed_branch_length


def weighted_unifrac_pair(u, v, u_total, v_total, branch_lengths):
    pu = u / u_total if u_total > 0 else u
    pv = v / v_total if v_total > 0 else v
    wu = (branch_lengths * np.absolute(pu - pv)).sum()
    return wu, pu, pv


def branch_correction(node_to_root_distances, pu, pv):
    return (node_to_root_distances.ravel() * (pu + pv)).sum()


def weighted_unifrac_normalized_pair(u, v, u_total, v_total, branch_lengths, d):
    if u_total == 0.0 and v_total == 0.0:
        return 0.0
    wu, pu, pv = weighted_unifrac_pair(u, v, u_total, v_total, branch_lengths)
    c = branch_correction(d, pu, pv)
    with np.errstate(divide="ignore", invalid="ignore"):
        return wu / c


def _pdist_callable(X, f):
    """SciPy's ``_pdist_callable`` loop (scipy/spatial/distance.py): condensed, row-major
    upper triangle."""
    n = X.shape[0]
    out = np.empty(n * (n - 1) // 2, dtype=np.float64)
    k = 0
    for i in range(n - 1):
        for j in range(i + 1, n):
            out[k] = f(X[i], X[j])
            k += 1
    return out


def unifrac_condensed(metric, counts, taxa, parent, length, names, normalized=False,
                      use_scipy=False):
    """Condensed distance vector exactly as ``beta_diversity(metric, counts, tree=...,
    taxa=...)`` produces it (_driver.py:293-355) for the two UniFrac metric names.

    ``counts`` is the raw table (``n_samples x n_taxa``).  For ``unweighted_unifrac`` the
    driver first applies ``counts > 0.0`` (_driver.py:303-304, _util.py:184-185).
    """
    counts = np.atleast_2d(np.asarray(counts))
    child_index = child_index_from_parent(parent)
    if metric == "unweighted_unifrac":
        counts = counts > 0.0
    a = nodes_by_counts(counts, taxa, names, child_index)
    X = a.T  # (n_samples, n_nodes), non-contiguous like the reference's
    bl = np.asarray(length, dtype=np.float64)
    if metric == "unweighted_unifrac":
        f = partial(unweighted_unifrac_pair, branch_lengths=bl)
    elif metric == "weighted_unifrac":
        tips = tip_indices_from_parent(parent)
        if normalized:
            d = tip_distances(bl, parent, tips)

            def f(u, v):
                ut = np.take(u, tips).sum()
                vt = np.take(v, tips).sum()
                return weighted_unifrac_normalized_pair(u, v, ut, vt, bl, d)
        else:

            def f(u, v):
                ut = np.take(u, tips).sum()
                vt = np.take(v, tips).sum()
                return weighted_unifrac_pair(u, v, ut, vt, bl)[0]
    else:
        raise ValueError(metric)
    if use_scipy:
        from scipy.spatial.distance import pdist

        return pdist(X, metric=f)
    return _pdist_callable(X, f)


# --------------------------------------------------------------------------------------
# SciPy metrics restated (the reference delegates these to scipy.spatial.distance.pdist)
# --------------------------------------------------------------------------------------
def braycurtis_condensed(counts):
    """``sum|u-v| / sum|u+v|`` in float64; 0/0 gives NaN as SciPy's compiled loop does."""
    X = np.atleast_2d(np.asarray(counts)).astype(np.float64)

    def f(u, v):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.float64(np.abs(u - v).sum()) / np.float64(np.abs(u + v).sum())

    return _pdist_callable(X, f)


def jaccard_condensed(counts):
    """``#(u != v and (u or v)) / #(u or v)`` on ``counts > 0``; two empty rows give 0."""
    X = np.atleast_2d(np.asarray(counts)) > 0.0

    def f(u, v):
        nonzero = np.logical_or(u, v)
        unequal = np.logical_and(u != v, nonzero)
        den = np.float64(nonzero.sum())
        return np.float64(unequal.sum()) / den if den != 0 else 0.0

    return _pdist_callable(X, f)


def faith_pd(counts, taxa, parent, length, names):
    """Faith's PD per sample (/root/reference/skbio/diversity/alpha/_pd.py:19-51):
    ``(branch_lengths * (counts_by_node > 0)).sum()``."""
    counts = np.atleast_2d(np.asarray(counts))
    a = nodes_by_counts(counts, taxa, names, child_index_from_parent(parent))
    bl = np.asarray(length, dtype=np.float64)
    return np.array([(bl * (a[:, s] > 0)).sum() for s in range(a.shape[1])])
