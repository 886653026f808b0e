"""Synthetic inputs of the shapes BASELINE.json names.

* ``random_join_tree``  - random binary tree by repeated uniform random pairwise joins of a
  pool of tips ``T0..T{tips-1}``; non-root branch lengths ~ U(0.01, 1.0); root length None.
* ``caterpillar_tree``  - ladder ``node_i = (node_{i-1}, S_i)`` with small random-join
  subtrees ``S_i`` of size ``1 + Geometric(0.5)`` (depth grows linearly with tips).
* ``poisson_csr``       - iid Poisson(lam) counts generated directly as CSR (non-zero
  positions Bernoulli(1 - exp(-lam)), values zero-truncated Poisson), int64.

Trees are produced as flat arrays in the reference's node id order
(``parent[k] > k``, siblings consecutive; /root/reference/skbio/tree/_tree.py:5337-5359) so
that large configurations never build a million Python node objects; ``to_newick`` /
``to_nodes`` materialise a tree object for the small cases that go through the full API.
"""

from __future__ import annotations

import numpy as np

from ._tree import SimpleTreeNode


def _ids_from_children(root, left, right):
    """Reference id order for a binary tree given child arrays (internal node -> children).

    Nodes are numbered in construction order on input (tips first).  Output:
    ``new_id[old]`` following assign_ids: children of a node get consecutive ids when the
    node is reached in postorder; root last.
    """
    n_total = len(left)
    new_id = np.full(n_total, -1, dtype=np.int64)
    counter = 0
    # iterative postorder
    stack = [(root, 0)]
    while stack:
        node, state = stack.pop()
        l = left[node]
        if l < 0:
            continue
        if state == 0:
            stack.append((node, 1))
            stack.append((right[node], 0))
            stack.append((l, 0))
        else:
            new_id[l] = counter
            new_id[right[node]] = counter + 1
            counter += 2
    new_id[root] = counter
    return new_id


def _finish(root, left, right, n_tips, rng, names_prefix="T"):
    n_total = len(left)
    new_id = _ids_from_children(root, left, right)
    parent_old = np.full(n_total, -1, dtype=np.int64)
    internal = np.nonzero(left >= 0)[0]
    parent_old[left[internal]] = internal
    parent_old[right[internal]] = internal
    parent = np.full(n_total, -1, dtype=np.int32)
    nz = parent_old >= 0
    parent[new_id[nz]] = new_id[parent_old[nz]].astype(np.int32)
    length = rng.uniform(0.01, 1.0, size=n_total)
    length[new_id[root]] = 0.0  # root length None -> 0.0
    names = np.empty(n_total, dtype=object)
    tip_old = np.arange(n_tips)
    tip_names = np.array([f"{names_prefix}{i}" for i in range(n_tips)], dtype=object)
    names[new_id[tip_old]] = tip_names
    tip_node = new_id[tip_old].astype(np.int32)  # node id of tip T_i
    return parent, length, names, tip_node


def random_join_tree(n_tips, seed=1):
    """Flat arrays ``(parent, length, names, tip_node)`` of a random-join binary tree."""
    rng = np.random.default_rng(seed)
    n_total = 2 * n_tips - 1
    left = np.full(n_total, -1, dtype=np.int64)
    right = np.full(n_total, -1, dtype=np.int64)
    pool = list(range(n_tips))
    nxt = n_tips
    # random pairwise joins: pick two distinct pool members uniformly
    rnd = rng.random(size=2 * max(n_tips - 1, 1))
    ri = 0
    while len(pool) > 1:
        i = int(rnd[ri] * len(pool)); ri += 1
        a = pool[i]
        pool[i] = pool[-1]
        pool.pop()
        j = int(rnd[ri] * len(pool)); ri += 1
        b = pool[j]
        left[nxt] = a
        right[nxt] = b
        pool[j] = nxt
        nxt += 1
    root = pool[0]
    return _finish(root, left, right, n_tips, rng)


def caterpillar_tree(n_tips, seed=1):
    """Deep, unbalanced ladder tree (BASELINE config 5).  Same return as random_join_tree."""
    rng = np.random.default_rng(seed)
    n_total = 2 * n_tips - 1
    left = np.full(n_total, -1, dtype=np.int64)
    right = np.full(n_total, -1, dtype=np.int64)
    nxt = n_tips
    used = 0

    def small_subtree(size):
        nonlocal nxt, used
        pool = list(range(used, used + size))
        used += size
        while len(pool) > 1:
            i = int(rng.integers(len(pool)))
            a = pool[i]; pool[i] = pool[-1]; pool.pop()
            j = int(rng.integers(len(pool)))
            b = pool[j]
            left[nxt] = a; right[nxt] = b
            pool[j] = nxt
            nxt += 1
        return pool[0]

    sizes = rng.geometric(0.5, size=n_tips) + 1
    si = 0
    first = min(int(sizes[si]), n_tips); si += 1
    spine = small_subtree(first)
    while used < n_tips:
        size = min(int(sizes[si]), n_tips - used); si += 1
        sub = small_subtree(size)
        left[nxt] = spine; right[nxt] = sub
        spine = nxt
        nxt += 1
    return _finish(spine, left, right, n_tips, rng)


def _zero_truncated_poisson(rng, lam, size):
    """Poisson(lam) conditioned on being >= 1, by inversion of the CDF."""
    kmax = int(lam + 12.0 * np.sqrt(lam + 1.0) + 40)
    k = np.arange(kmax + 1)
    logpmf = -lam + k * np.log(lam) - np.cumsum(np.log(np.maximum(k, 1)))
    cdf = np.cumsum(np.exp(logpmf))
    u = cdf[0] + (1.0 - cdf[0]) * rng.random(size)
    return np.minimum(np.searchsorted(cdf, u, side="right"), kmax).astype(np.int64)


def poisson_csr(n_samples, n_cols, lam, seed=0, chunk_cells=1 << 27):
    """iid Poisson(lam) counts as CSR ``(indptr int64, indices int32, values int64)``.

    Column indices within a row are ascending.  The non-zero cells are iid Bernoulli
    (1 - exp(-lam)): their positions in the flattened table are generated as cumulative sums
    of Geometric(p) gaps (fully vectorised, no dense table); values are zero-truncated
    Poisson(lam) by rejection.
    """
    rng = np.random.default_rng(seed)
    p = 1.0 - np.exp(-lam)
    rows_per_chunk = max(1, int(chunk_cells // max(n_cols, 1)))
    counts = np.zeros(n_samples, dtype=np.int64)
    idx_parts, val_parts = [], []
    for s0 in range(0, n_samples, rows_per_chunk):
        s1 = min(n_samples, s0 + rows_per_chunk)
        cells = (s1 - s0) * n_cols
        pos_parts, last = [], -1
        while last < cells - 1:
            remaining = cells - 1 - last
            k = int(remaining * p + 6.0 * np.sqrt(remaining * p + 1.0) + 16)
            gaps = rng.geometric(p, size=k)
            pos = last + np.cumsum(gaps)
            pos = pos[pos < cells]
            pos_parts.append(pos)
            if len(pos) < k:
                break
            last = int(pos[-1])
        pos = np.concatenate(pos_parts) if pos_parts else np.zeros(0, dtype=np.int64)
        r = pos // n_cols
        c = pos - r * n_cols
        counts[s0:s1] = np.bincount(r, minlength=s1 - s0)
        vals = _zero_truncated_poisson(rng, lam, len(pos))
        idx_parts.append(c.astype(np.int32))
        val_parts.append(vals.astype(np.int64))
    indptr = np.zeros(n_samples + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = np.concatenate(idx_parts) if idx_parts else np.zeros(0, dtype=np.int32)
    values = np.concatenate(val_parts) if val_parts else np.zeros(0, dtype=np.int64)
    return indptr, indices, values


def csr_to_dense(indptr, indices, values, n_cols):
    n = len(indptr) - 1
    out = np.zeros((n, n_cols), dtype=values.dtype)
    rows = np.repeat(np.arange(n), np.diff(indptr))
    out[rows, indices] = values
    return out


def to_nodes(parent, length, names, root_length_none=True):
    """Materialise ``SimpleTreeNode`` objects from flat arrays (small trees only)."""
    m = len(parent)
    nodes = [SimpleTreeNode(name=names[k], length=float(length[k])) for k in range(m)]
    root = None
    for k in range(m):  # ascending id keeps sibling order
        p = parent[k]
        if p < 0:
            root = nodes[k]
        else:
            nodes[p].append(nodes[k])
    if root_length_none and root is not None:
        root.length = None
    return root


def to_newick(parent, length, names, root_length_none=True):
    """Newick text for flat arrays; lengths written with ``repr`` so they round-trip."""
    m = len(parent)
    children = [[] for _ in range(m)]
    root = -1
    for k in range(m):
        p = parent[k]
        if p < 0:
            root = k
        else:
            children[p].append(k)
    out = []
    stack = [(root, 0)]
    while stack:
        node, i = stack.pop()
        ch = children[node]
        if i == 0 and ch:
            out.append("(")
        if i < len(ch):
            if i > 0:
                out.append(",")
            stack.append((node, i + 1))
            stack.append((ch[i], 0))
        else:
            if ch:
                out.append(")")
            nm = names[node]
            if nm is not None:
                out.append("'" + str(nm).replace("'", "''") + "'")
            if not (node == root and root_length_none):
                out.append(":" + repr(float(length[node])))
    out.append(";")
    return "".join(out)
