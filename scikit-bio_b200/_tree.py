"""Host-side tree handling: flatten a TreeNode into flat arrays, validate taxa.

The engine never walks linked node objects on the hot path.  This module does the one
traversal that turns a ``skbio.TreeNode`` (or any duck-typed node exposing ``children``,
``name``, ``length``) into

* ``parent[m]``  int32, index of each node's parent, ``-1`` for the root,
* ``length[m]``  float64, branch length, ``None``/NaN replaced by ``0.0``,
* ``name[m]``    object, node name or ``None``,

with nodes numbered exactly as the reference numbers them (``TreeNode.assign_ids``,
/root/reference/skbio/tree/_tree.py:5337-5359: the children of each node receive consecutive
ids when that node is visited in postorder, the root gets the last id).  That numbering has
the property every later stage relies on: ``parent[k] > k`` for every non-root node.

Reference behaviour mirrored here:

* ``TreeNode.to_array(nan_length_value=0.0)``  /root/reference/skbio/tree/_tree.py:6235-6329
* ``_validate_taxa_and_tree(rooted=True, lengths=True)``  /root/reference/skbio/tree/_utils.py:29-90
* name -> node lookup of ``_nodes_by_counts``  /root/reference/skbio/diversity/_phylogenetic.pyx:190-206
"""

from __future__ import annotations

import numpy as np

try:  # use the reference's exception types when scikit-bio itself is importable
    from skbio.tree import DuplicateNodeError, MissingNodeError  # type: ignore
except Exception:  # pragma: no cover - exercised on boxes without scikit-bio

    class TreeError(Exception):
        """General tree error (mirror of skbio.tree.TreeError)."""

    class DuplicateNodeError(TreeError):
        """Duplicate nodes with identical names."""

    class MissingNodeError(TreeError):
        """Expecting a node."""


class SimpleTreeNode:
    """Minimal stand-in for ``skbio.TreeNode`` (name, length, parent, children, id).

    Only what the beta-diversity path touches is implemented.  It exists so that tests,
    the benchmark and users without scikit-bio installed can still hand a tree object to
    ``beta_diversity``; a real ``skbio.TreeNode`` is accepted everywhere this is.
    """

    __slots__ = ("name", "length", "parent", "children", "id")

    def __init__(self, name=None, length=None, children=None):
        self.name = name
        self.length = length
        self.parent = None
        self.children = []
        self.id = None
        if children:
            for c in children:
                self.append(c)

    def append(self, node):
        node.parent = self
        self.children.append(node)

    def is_tip(self):
        return not self.children

    def is_root(self):
        return self.parent is None

    def root(self):
        n = self
        while n.parent is not None:
            n = n.parent
        return n

    def postorder(self, include_self=True):
        yield from _postorder(self, include_self)

    def tips(self):
        for n in _postorder(self, True):
            if not n.children:
                yield n

    def __bool__(self):  # a node is always truthy (reference: bool(node) == has children)
        return True

    @classmethod
    def read(cls, newick):
        return read_newick(newick)


def _postorder(root, include_self=True):
    """Iterative postorder over duck-typed nodes (children before parents, left to right)."""
    stack = [(root, 0)]
    while stack:
        node, i = stack.pop()
        ch = node.children
        if i < len(ch):
            stack.append((node, i + 1))
            stack.append((ch[i], 0))
        else:
            if node is root and not include_self:
                continue
            yield node


def read_newick(text):
    """Parse a Newick string into ``SimpleTreeNode`` objects.

    Supports quoted labels, ``[...]`` comments and ``:length``; underscores in unquoted
    labels become spaces as in the reference reader
    (/root/reference/skbio/io/format/newick.py:270-336 - behaviour only, independent code).
    """
    if hasattr(text, "read"):
        text = text.read()
    s = text
    n = len(s)
    i = 0
    root = SimpleTreeNode()

    def skip_ws_comments(i):
        while i < n:
            c = s[i]
            if c.isspace():
                i += 1
            elif c == "[":
                depth = 1
                i += 1
                while i < n and depth:
                    if s[i] == "[":
                        depth += 1
                    elif s[i] == "]":
                        depth -= 1
                    i += 1
            else:
                break
        return i

    stack = []
    node = root
    # ``node`` is the node currently being described (its label follows)
    while True:
        i = skip_ws_comments(i)
        if i >= n:
            break
        c = s[i]
        if c == "(":
            child = SimpleTreeNode()
            node.append(child)
            stack.append(node)
            node = child
            i += 1
        elif c == ",":
            parent = stack[-1]
            sib = SimpleTreeNode()
            parent.append(sib)
            node = sib
            i += 1
        elif c == ")":
            node = stack.pop()
            i += 1
        elif c == ";":
            break
        elif c == ":":
            i += 1
            i = skip_ws_comments(i)
            j = i
            while j < n and s[j] not in ",();[ \t\r\n":
                j += 1
            node.length = float(s[i:j])
            i = j
        elif c == "'":
            j = i + 1
            buf = []
            while j < n:
                if s[j] == "'":
                    if j + 1 < n and s[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(s[j])
                j += 1
            node.name = "".join(buf)
            i = j + 1
        else:
            j = i
            while j < n and s[j] not in ":,();[":
                j += 1
            label = s[i:j].strip().replace("_", " ")
            node.name = label if label else None
            i = j
    if stack:
        raise ValueError("Unbalanced parentheses in Newick string.")
    return root


class FlatTree:
    """A tree as flat arrays in the reference's node id order (root last, ``parent[k] > k``).

    Accepted wherever ``beta_diversity`` takes ``tree=``; skips node objects altogether.
    ``length`` holds NaN where the source gave no branch length (``None`` on a TreeNode).
    Build one with ``read_newick_flat`` (native parser: the names then stay one byte buffer,
    ``name_blob``, and ``names`` is only decoded if somebody asks) or from your own arrays.
    """

    __slots__ = ("parent", "length", "_names", "name_blob")

    def __init__(self, parent, length, names=None, name_blob=None):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.length = np.ascontiguousarray(length, dtype=np.float64)
        self.name_blob = name_blob
        self._names = None if names is None else np.asarray(names, dtype=object)
        if names is None and name_blob is None:
            raise ValueError("FlatTree needs names (or a name_blob)")
        m = len(self._names) if self._names is not None else len(name_blob.has)
        if not (len(self.parent) == len(self.length) == m):
            raise ValueError("parent, length and names must have the same length")

    @property
    def names(self):
        if self._names is None:
            self._names = self.name_blob.decode()
        return self._names

    def __len__(self):
        return len(self.parent)

    def tip_mask(self):
        has_child = np.zeros(len(self.parent), dtype=bool)
        has_child[self.parent[self.parent >= 0]] = True
        return ~has_child


def read_newick_flat(text):
    """Newick text (or an open file) -> ``FlatTree``, parsed natively (libskbb200) straight into
    reference id order."""
    from . import _capi

    if hasattr(text, "read"):
        text = text.read()
    parent, length, blob = _capi.newick_to_flat(text)
    return FlatTree(parent, length, name_blob=blob)


def _root_of(tree):
    r = getattr(tree, "root", None)
    if callable(r):
        return r()
    node = tree
    while getattr(node, "parent", None) is not None:
        node = node.parent
    return node


def flatten_for_engine(tree):
    """``flatten_tree`` for the engine's host path: the names of a natively parsed ``FlatTree`` stay
    a ``NameBlob`` (no per-node Python strings are ever made)."""
    if isinstance(tree, FlatTree) and tree._names is None:
        length = tree.length.copy()
        length[np.isnan(length)] = 0.0
        return tree.parent, length, tree.name_blob
    return flatten_tree(tree)


def flatten_tree(tree, assign_ids=True):
    """Flatten ``tree`` (the node passed in is treated as the root, as the reference does).

    Returns ``(parent, length, names)`` in reference id order.  One traversal instead of the
    reference's three (``assign_ids`` + ``index_tree`` + ``traverse`` in ``to_array``).

    As in the reference, ``node.id`` is written on every node when ``assign_ids`` is true
    (/root/reference/skbio/tree/_tree.py:5353-5359).
    """
    if isinstance(tree, FlatTree):
        length = tree.length.copy()
        length[np.isnan(length)] = 0.0  # to_array(nan_length_value=0.0)
        return tree.parent, length, tree.names
    nodes = []  # nodes in id order
    parents = []  # parent *object index into ``order``* resolved afterwards
    # ids are handed out to the children of each node when the node is reached in postorder
    idx_of = {}
    par_list = []
    for node in _postorder(tree, True):
        ch = node.children
        if ch:
            first = len(nodes)
            nodes.extend(ch)
            par_list.append((node, first, len(ch)))
    nodes.append(tree)
    m = len(nodes)
    parent = np.full(m, -1, dtype=np.int32)
    # a node's own id is only known once *its* parent was visited, so resolve in a second,
    # array-only pass: record object identity -> id.
    for k, nd in enumerate(nodes):
        idx_of[id(nd)] = k
    for node, first, cnt in par_list:
        parent[first : first + cnt] = idx_of[id(node)]
    length = np.empty(m, dtype=np.float64)
    names = np.empty(m, dtype=object)
    for k, nd in enumerate(nodes):
        ln = nd.length
        length[k] = np.nan if ln is None else ln
        names[k] = nd.name
        if assign_ids:
            try:
                nd.id = k
            except Exception:  # read-only foreign node types
                assign_ids = False
    length[np.isnan(length)] = 0.0  # to_array(nan_length_value=0.0), _tree.py:6326-6328
    return parent, length, names


def validate_taxa_and_tree(taxa, tree, rooted=True, _qblob=None):
    """Same checks, order, exception types and messages as the reference's
    ``_validate_taxa_and_tree(taxa, tree, rooted=True, lengths=True)``
    (/root/reference/skbio/tree/_utils.py:29-90)."""
    taxon_set = set(taxa)
    if len(taxa) != len(taxon_set):
        raise ValueError("All taxa must be unique.")
    if isinstance(tree, FlatTree):
        m = len(tree)
        if rooted and int(np.count_nonzero(tree.parent == m - 1)) > 2:
            raise ValueError("The tree must be rooted.")
        if np.isnan(tree.length[: m - 1]).any():
            raise ValueError("All non-root nodes in the tree must have a branch length.")
        if tree._names is None:   # names still one byte buffer: the name checks run natively
            from . import _capi

            dup, n_missing = _capi.names_validate(tree.name_blob, tree.parent, list(taxa), _qblob)
        else:
            tip_names = list(tree.names[:m][tree.tip_mask()]) if m > 1 else []
            tip_name_set = set(tip_names)
            dup = len(tip_names) != len(tip_name_set)
            n_missing = len(taxon_set - tip_name_set)
        if dup:
            raise DuplicateNodeError("All tip names in the tree must be unique.")
        if n_missing:
            raise MissingNodeError(
                f"{n_missing} taxa are not present as tip names in the tree.")
        return
    if rooted and len(_root_of(tree).children) > 2:
        raise ValueError("The tree must be rooted.")
    tip_names = []
    for node in _postorder(tree, include_self=False):
        if node.length is None:
            raise ValueError("All non-root nodes in the tree must have a branch length.")
        if not node.children:
            tip_names.append(node.name)
    tip_name_set = set(tip_names)
    if len(tip_names) != len(tip_name_set):
        raise DuplicateNodeError("All tip names in the tree must be unique.")
    missing = taxon_set - tip_name_set
    if missing:
        raise MissingNodeError(
            f"{len(missing)} taxa are not present as tip names in the tree."
        )


def map_all_taxa(taxa, names, _qblob=None):
    """Node id of every taxon, ``-1`` where no node carries the name.

    Mirrors the lookup in ``_nodes_by_counts`` (_phylogenetic.pyx:190-206): all node names are
    scanned in id order, tips and internal nodes alike, and a later node with the same name
    overwrites an earlier one.  `names` is an object array or a ``NameBlob`` (native lookup)."""
    from . import _capi

    taxa_list = taxa.tolist() if isinstance(taxa, np.ndarray) else list(taxa)
    if isinstance(names, _capi.NameBlob):
        return _capi.names_lookup(names, taxa_list, _qblob)
    lookup = dict(zip(names.tolist() if isinstance(names, np.ndarray) else list(names), range(len(names))))
    lookup.pop(None, None)
    return np.fromiter((lookup.get(t, -1) for t in taxa_list), dtype=np.int32, count=len(taxa_list))


def map_taxa_to_nodes(taxa, observed_cols, names):
    """Node id for every *observed* taxon column; ``-1`` for unobserved columns.  An observed
    taxon that matches no node raises ``KeyError`` exactly like the reference's
    ``node_lookup[n]`` does (only reachable with ``validate=False``)."""
    node_of_col = map_all_taxa(taxa, names)
    out = np.full(len(node_of_col), -1, dtype=np.int32)
    observed_cols = np.asarray(observed_cols, dtype=np.int64)
    out[observed_cols] = node_of_col[observed_cols]
    bad = observed_cols[out[observed_cols] < 0]
    if len(bad):
        raise KeyError(taxa[int(bad[0])])
    return out
