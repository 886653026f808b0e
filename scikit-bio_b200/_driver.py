"""``beta_diversity`` - drop-in for ``skbio.diversity.beta_diversity`` on the B200 engine.

Same signature, validation order, exception types/messages and return type as
/root/reference/skbio/diversity/_driver.py:231-355 for

    unweighted_unifrac, weighted_unifrac (normalized=False|True), braycurtis, jaccard

which run on the GPU through libskbb200.so (no CPU path: a missing library or GPU raises
``EngineError``).  Every other metric / a user ``pairwise_func`` / a callable metric is not on
this engine's path and is handed to the reference's own route (scikit-bio if importable,
else the same ``scipy.spatial.distance.pdist`` call the reference makes).

Engine-only keyword arguments (not forwarded to any metric):
    devices   : int or list of int, GPUs to shard the pair triangle over (default [0])
    condensed : keep the result in condensed form (skbio >= 0.7 ``DistanceMatrix(...,
                condensed=True)``).  Default: square like the reference up to 16,384 samples,
                condensed above (the square form of 100,000 samples is 80 GB)
"""

from __future__ import annotations

import threading
import time
import warnings
import weakref
from functools import partial
from inspect import signature

import numpy as np

from . import _capi
from ._dm import make_distance_matrix
from ._table import ingest_table, to_csr, validate_counts_matrix, _is_sparse
from ._tree import (FlatTree, flatten_for_engine, flatten_tree, map_all_taxa, map_taxa_to_nodes,
                    validate_taxa_and_tree)

# /root/reference/skbio/diversity/_driver.py:43-52
_qualitative_metrics = {
    "dice", "jaccard", "matching", "rogerstanimoto", "russellrao", "sokalsneath", "yule",
    "unweighted_unifrac",
}
_slow_beta_callables = {"unweighted_unifrac", "weighted_unifrac"}
# /root/reference/skbio/diversity/_driver.py:58-84
_pdist_metrics = {
    "euclidean", "cityblock", "braycurtis", "canberra", "chebyshev", "correlation", "cosine",
    "dice", "hamming", "jaccard", "jensenshannon", "mahalanobis", "manhattan", "matching",
    "minkowski", "rogerstanimoto", "russellrao", "seuclidean", "sokalsneath", "sqeuclidean",
    "yule",
}
_engine_metrics = {"unweighted_unifrac", "weighted_unifrac", "braycurtis", "jaccard"}
_normalize_weighted_unifrac_by_default = False  # beta/_unifrac.py:26


def _get_phylogenetic_kwargs(kwargs, taxa):
    # /root/reference/skbio/diversity/_util.py:173-181
    if taxa is None:
        raise ValueError("`taxa` is required for phylogenetic diversity metrics.")
    try:
        tree = kwargs.pop("tree")
    except KeyError:
        raise ValueError("`tree` is required for phylogenetic diversity metrics.")
    return taxa, tree, kwargs


def _devices_arg(devices):
    if devices is None:
        return [0]
    if isinstance(devices, (int, np.integer)):
        return [int(devices)]
    devs = [int(d) for d in devices]
    if not devs:
        raise ValueError("`devices` must name at least one GPU.")
    return devs


# condensed results above this many samples stay condensed unless the caller asks otherwise
# (the reference always expands to n x n, _base.py:1292-1299: +80 GB at 100,000 samples)
_AUTO_CONDENSED_ABOVE = 16384

_profile = {}


def last_call_profile():
    """Host-side time split (seconds) of the last engine call on this thread's process:
    validate / flatten / csr / lookup / engine (create + compute + fetch) / result object."""
    return dict(_profile)


_pinned_cache = []          # at most one idle page-locked result buffer (reused by the next call)
_PINNED_RESULT_MAX = 8 << 30  # larger results go to pageable memory (page-locking them costs more than it saves)


def _result_vector(total):
    """The condensed result vector.  Up to 8 GB it lives in page-locked memory (the D2H copy is
    then plain DMA, ~50 GB/s instead of ~7 GB/s through the driver's bounce buffers); the buffer
    goes back to a one-entry cache when the last array referring to it is garbage-collected, so
    a loop of calls page-locks once."""
    if total == 0 or total * 8 > _PINNED_RESULT_MAX:
        return np.empty(total, dtype=np.float64)
    buf = None
    while _pinned_cache:
        cand = _pinned_cache.pop()
        if cand.n >= total and cand.n <= 2 * total + 1024:
            buf = cand
            break
        cand.free()
    if buf is None:
        buf = _capi.PinnedBuffer(total)
    buf.array = None                 # the buffer object must not keep the array alive itself
    whole = buf.new_array()

    def _release(b=buf):
        if _pinned_cache:
            b.free()
        else:
            _pinned_cache.append(b)

    weakref.finalize(whole, _release)   # every view (the slice below, the DistanceMatrix' data) keeps `whole` alive
    return whole[:total]


def _run_sharded(make_problem, metric_key, n, devices):
    """One problem (shard) per device, each driven by its own host thread; every shard writes
    its own disjoint slices of ONE condensed vector (no gather step, no extra copy).  With
    several shards the per-block products are built once - each shard builds its own block range
    and the shards exchange them by peer copies (distributed.ThreadCollectives)."""
    if _capi.load().b2d_device_count() <= 0:
        raise _capi.EngineError(
            "no CUDA device visible: the B200 engine has no CPU fallback for "
            f"'{metric_key}'.")
    out = _result_vector(n * (n - 1) // 2)
    errors = []
    coll = None
    if len(devices) > 1:
        from .distributed import ThreadCollectives
        coll = ThreadCollectives(len(devices))

    def work(shard, dev):
        try:
            with make_problem(dev, shard, len(devices)) as prob:
                if coll is not None:
                    prob.set_collectives(shard, len(devices), coll.for_rank(shard))
                prob.compute(metric_key)
                prob.fetch(out)
        except BaseException as e:  # propagate to the caller thread
            errors.append(e)
            if coll is not None:
                coll.abort()        # the other shards must not wait for this one

    if len(devices) == 1:
        work(0, devices[0])
    else:
        threads = [threading.Thread(target=work, args=(s, d)) for s, d in enumerate(devices)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        first = [e for e in errors if "BrokenBarrier" not in repr(e)] or errors
        raise first[0]
    return out


def _unifrac_engine(metric_key, counts, taxa, tree, validate, devices):
    """Replaces _setup_multiple_{un,}weighted_unifrac + pdist
    (/root/reference/skbio/diversity/beta/_unifrac.py:474-584, _driver.py:349-354).  Several
    observed taxa mapping to one node (only reachable with validate=False) follow the
    reference's assignment semantics: the last column wins (_phylogenetic.pyx:208-218)."""
    make = _prepare_unifrac_problem(metric_key, counts, taxa, tree, validate)
    t0 = time.perf_counter()
    vec = _run_sharded(make, metric_key, counts.shape[0], devices)
    _profile["engine_s"] = time.perf_counter() - t0
    return vec


def _dedupe_last_wins(indptr, cols, node_ids, vals):
    """Several observed columns mapped to one node: the reference assigns
    ``count_array[node, :] = counts[:, col]`` column by column in ascending order
    (_phylogenetic.pyx:208-218), zeros included, so the LAST observed column of a node decides
    every sample's count.  Keep only the CSR entries of those winning columns."""
    n = len(indptr) - 1
    winner = np.full(int(node_ids.max()) + 1 if len(node_ids) else 1, -1, dtype=np.int64)
    np.maximum.at(winner, node_ids, cols.astype(np.int64))
    keep = winner[node_ids] == cols
    rows = np.repeat(np.arange(n), np.diff(indptr))[keep]
    new_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=new_ptr[1:])
    v = None if vals is None else np.ascontiguousarray(vals[keep])
    return new_ptr, np.ascontiguousarray(node_ids[keep], dtype=np.int32), v


def _features_engine(metric_key, counts, devices):
    mode = "presence" if metric_key == "jaccard" else "float64"
    indptr, cols, vals = to_csr(counts, mode)
    n, f = counts.shape

    def make(dev, shard, nshards):
        return _capi.Problem.features(indptr, cols, vals, n, f, device=dev, shard=shard,
                                      n_shards=nshards)

    return _run_sharded(make, metric_key, n, devices)


def beta_diversity(metric, counts, ids=None, validate=True, pairwise_func=None, **kwargs):
    r"""Compute distances between all pairs of samples (see the module docstring).

    Parameters and behaviour follow ``skbio.diversity.beta_diversity``
    (/root/reference/skbio/diversity/_driver.py:231-355).
    """
    devices = _devices_arg(kwargs.pop("devices", None))
    condensed = kwargs.pop("condensed", None)
    taxa = kwargs.pop("taxa", None)
    t_start = time.perf_counter()
    counts, ids, taxa = ingest_table(counts, ids, taxa)
    if 0 in counts.shape:
        # _driver.py:295-300
        return make_distance_matrix(np.zeros((len(ids), len(ids))), ids)
    if validate:
        counts = validate_counts_matrix(counts)

    on_engine = isinstance(metric, str) and metric in _engine_metrics and pairwise_func is None
    if not on_engine:
        return _reference_route(metric, counts, ids, validate, pairwise_func, taxa, kwargs)

    if ids is not None:
        ids = tuple(ids)
    n = counts.shape[0]
    if metric in ("unweighted_unifrac", "weighted_unifrac"):
        taxa, tree, kwargs = _get_phylogenetic_kwargs(kwargs, taxa)
        key = metric
        if metric == "weighted_unifrac":
            normalized = kwargs.pop("normalized", _normalize_weighted_unifrac_by_default)
            if normalized:
                key = "weighted_unifrac_normalized"
        if kwargs:
            # the reference forwards leftovers to the pair function, which rejects them
            # (test_driver.py:412-428)
            raise TypeError(
                f"{metric}() got an unexpected keyword argument '{next(iter(kwargs))}'")
        vec = _unifrac_engine(key, counts, taxa, tree, validate, devices)
    else:
        if kwargs:
            raise TypeError(
                f"pdist metric '{metric}' got an unexpected keyword argument "
                f"'{next(iter(kwargs))}'")
        vec = _features_engine(metric, counts, devices)
    if condensed is None:
        condensed = n > _AUTO_CONDENSED_ABOVE
    t_dm = time.perf_counter()
    dm = make_distance_matrix(vec, ids, condensed=bool(condensed), n=n)
    _profile["result_object_s"] = time.perf_counter() - t_dm
    _profile["total_s"] = time.perf_counter() - t_start
    return dm


def _unifrac_host_arrays(key, counts, taxa, tree, validate):
    """Host half of the UniFrac path: validate, flatten, CSR, name lookup.  Returns the flat
    arrays the C ABI takes: ``(parent, length, indptr, node_ids, values, n_samples)``."""
    h = _unifrac_host(key, counts, taxa, tree, validate)
    ids = h["ids"] if h["node_of_col"] is None else h["node_of_col"][h["ids"]]
    return h["parent"], h["length"], h["indptr"], ids, h["vals"], h["n"]


def _unifrac_host(key, counts, taxa, tree, validate):
    """validate -> flatten -> CSR -> taxon lookup.  The nnz-long index array is never rewritten on
    the host: in the normal case the CSR keeps its table COLUMNS and the device applies the column
    -> node table (``node_of_col``); only when several observed columns share a node (reachable
    with validate=False) is the table applied here, followed by the reference's last-column-wins
    rule."""
    t0 = time.perf_counter()
    # the taxon names as one byte buffer, built once for the native name checks and the lookup
    qblob = _capi.query_blob(taxa) if isinstance(tree, FlatTree) and tree._names is None else None
    if validate:
        validate_taxa_and_tree(taxa, tree, _qblob=qblob)
    t1 = time.perf_counter()
    parent, length, names = flatten_for_engine(tree)
    t2 = time.perf_counter()
    mode = "presence" if key == "unweighted_unifrac" else "int64"
    indptr, cols, vals = to_csr(counts, mode)
    t3 = time.perf_counter()
    n = counts.shape[0]
    node_of_col = map_all_taxa(taxa, names, qblob)
    mapped = node_of_col[node_of_col >= 0]
    out = dict(parent=parent, length=length, indptr=indptr, ids=cols, vals=vals, n=n,
               node_of_col=node_of_col)
    shared = len(mapped) and int(np.bincount(mapped, minlength=len(parent)).max()) > 1
    if shared or len(mapped) != len(node_of_col):
        # rare: duplicate nodes among the taxa, or taxa without a node - decide on the OBSERVED columns
        observed = np.zeros(counts.shape[1], dtype=bool)
        observed[cols] = True
        observed_cols = np.flatnonzero(observed)
        node_of_obs = map_taxa_to_nodes(taxa, observed_cols, names)   # KeyError for a missing one
        if len(np.unique(node_of_obs[observed_cols])) != len(observed_cols):
            indptr, ids, vals = _dedupe_last_wins(indptr, cols, node_of_obs[cols], vals)
            out.update(indptr=indptr, ids=ids, vals=vals, node_of_col=None)
        else:
            out.update(node_of_col=node_of_obs)
    t4 = time.perf_counter()
    _profile.clear()
    _profile.update(validate_s=t1 - t0, flatten_s=t2 - t1, csr_s=t3 - t2, lookup_s=t4 - t3)
    return out


def _prepare_unifrac_problem(key, counts, taxa, tree, validate):
    """Host arrays -> factory of shards (one problem per device)."""
    h = _unifrac_host(key, counts, taxa, tree, validate)

    def make(dev, shard=0, nshards=1):
        return _capi.Problem.tree(h["parent"], h["length"], h["indptr"], h["ids"], h["vals"], h["n"],
                                  device=dev, shard=shard, n_shards=nshards, node_of_col=h["node_of_col"])

    return make


def partial_beta_diversity(metric, counts, ids, id_pairs, validate=True, **kwargs):
    """Distances only between the given ID pairs; every other entry of the matrix is 0.

    Same contract as the reference (/root/reference/skbio/diversity/_driver.py:364-460): only
    the optimized UniFrac metric names and callables are accepted, ``id_pairs`` must be a
    subset of ``ids`` without duplicates or self pairs, the result is ``DistanceMatrix(dm +
    dm.T, ids)``.  On the engine only the 128x128 tiles that contain a requested pair run.
    """
    from itertools import chain

    device = _devices_arg(kwargs.pop("devices", None))[0]
    taxa = kwargs.pop("taxa", None)
    counts, ids, taxa = ingest_table(counts, ids, taxa)
    if validate:
        counts = validate_counts_matrix(counts)
    id_pairs = list(id_pairs)
    all_ids_in_pairs = set(chain.from_iterable(id_pairs))
    if not all_ids_in_pairs.issubset(ids):
        raise ValueError("`id_pairs` are not a subset of `ids`")
    hashes = {i for i in id_pairs}.union({i[::-1] for i in id_pairs})
    if len(hashes) != len(id_pairs) * 2:
        raise ValueError("A duplicate or a self-self pair was observed.")
    n = len(ids)
    id_index = {id_: idx for idx, id_ in enumerate(ids)}
    pairs = np.array([(id_index[u], id_index[v]) for u, v in id_pairs], dtype=np.int64).reshape(-1, 2)
    dm = np.zeros((n, n))
    if metric in ("unweighted_unifrac", "weighted_unifrac"):
        key = metric
        if metric == "weighted_unifrac":
            normalized = kwargs.pop("normalized", _normalize_weighted_unifrac_by_default)
            if normalized:
                key = "weighted_unifrac_normalized"
        taxa, tree, kwargs = _get_phylogenetic_kwargs(kwargs, taxa)
        if kwargs:
            raise TypeError(
                f"{metric}() got an unexpected keyword argument '{next(iter(kwargs))}'")
        if _capi.load().b2d_device_count() <= 0:
            raise _capi.EngineError("no CUDA device visible: the B200 engine has no CPU fallback.")
        make = _prepare_unifrac_problem(key, counts, taxa, tree, validate)
        if len(pairs):
            lo = pairs.min(axis=1)
            hi = pairs.max(axis=1)
            tiles = np.unique(np.stack([lo // 128, hi // 128], axis=1), axis=0)
            with make(device) as prob:
                prob.select_tiles(tiles)
                prob.compute(key)
                compact, layout = prob.fetch_compact()
            # position of pair (i<j) inside the compact stripes
            stripe_first = {}
            pos = 0
            for first, cnt in layout:
                stripe_first[int(first)] = pos
                pos += int(cnt)
            b = lo // 128
            first_of_b = (b * 128) * n - ((b * 128) * (b * 128 + 1)) // 2
            cidx = lo * n - ((lo + 2) * (lo + 1)) // 2 + hi
            base = np.array([stripe_first[int(f)] for f in first_of_b], dtype=np.int64)
            vals = compact[base + (cidx - first_of_b)]
            dm[pairs[:, 0], pairs[:, 1]] = vals
    elif callable(metric):
        func = partial(metric, **kwargs)
        dense = counts.toarray() if _is_sparse(counts) else counts
        for u, v in pairs:
            dm[u, v] = func(dense[u], dense[v])
    else:
        raise ValueError(
            "partial_beta_diversity is only compatible with "
            "optimized unifrac methods and callable functions.")
    return make_distance_matrix(dm + dm.T, ids)


def block_beta_diversity(metric, counts, ids=None, validate=True, k=64, reduce_f=None,
                         map_f=None, **kwargs):
    """Block-decomposition beta diversity (/root/reference/skbio/diversity/_block.py:263-356).

    The reference cuts the triangle into k x k blocks so that a user-supplied ``map_f`` can
    farm them out; the result equals ``beta_diversity`` (its own test_block.py asserts that).
    The engine already computes in 128 x 128 tiles sharded over GPUs, so without a custom
    ``map_f``/``reduce_f`` this is the tiled engine call itself.  With one, the reference's
    structure is kept: blocks of ``k`` ids are mapped through ``partial_beta_diversity``.
    """
    if map_f is None and reduce_f is None:
        return beta_diversity(metric, counts, ids, validate=validate, **kwargs)
    taxa = kwargs.get("taxa")
    counts, ids, taxa = ingest_table(counts, ids, taxa)
    if taxa is not None:
        kwargs["taxa"] = taxa
    if validate:
        counts = validate_counts_matrix(counts)
    n = counts.shape[0]
    num_ids = np.arange(n)

    def block_jobs():
        for r0 in range(0, n, k):
            for c0 in range(r0, n, k):
                rids = num_ids[r0:r0 + k]
                cids = num_ids[c0:c0 + k]
                keep = np.unique(np.hstack([rids, cids]))
                if r0 == c0:
                    pairs = [(int(a), int(b)) for x, a in enumerate(rids) for b in rids[x + 1:]]
                else:
                    pairs = [(int(a), int(b)) for a in rids for b in cids]
                kw = dict(kwargs)
                kw.update(metric=metric, counts=counts[keep], ids=list(keep), id_pairs=pairs,
                          validate=False)
                yield kw

    def compute(**kw):
        return partial_beta_diversity(**kw)

    if map_f is None:
        def map_f(func, gen):
            for kw in gen:
                yield func(**kw)
    blocks = list(map_f(compute, block_jobs()))
    if reduce_f is not None:
        dm = reduce_f(blocks)
        try:
            dm.ids = ids
        except Exception:
            pass
        return dm
    mat = np.zeros((n, n))
    for blk in blocks:
        bid = np.asarray(blk.ids, dtype=np.int64)
        data = np.asarray(blk.data)
        mat[np.ix_(bid, bid)] += np.triu(data, k=1)
    return make_distance_matrix(mat + mat.T, ids)


def alpha_diversity(metric, counts, ids=None, validate=True, **kwargs):
    """``skbio.diversity.alpha_diversity`` for ``metric='faith_pd'`` on the engine
    (/root/reference/skbio/diversity/_driver.py:134-228, alpha/_pd.py:19-51): Faith's PD is the
    per-sample sum of branch lengths of observed nodes, a by-product of the presence
    propagation.  Other alpha metrics are not on this engine's path and go to scikit-bio."""
    import pandas as pd

    device = _devices_arg(kwargs.pop("devices", None))[0]
    if metric == "phydiv":
        return _phydiv_engine(counts, ids, validate, device, kwargs)
    if metric != "faith_pd":
        from skbio.diversity import alpha_diversity as _skbio_alpha  # type: ignore

        return _skbio_alpha(metric, counts, ids, validate=validate, **kwargs)
    taxa = kwargs.pop("taxa", None)
    counts, ids, taxa = ingest_table(counts, ids, taxa)
    if 0 in counts.shape:
        return pd.Series(np.zeros(len(ids)), index=ids)
    if validate:
        counts = validate_counts_matrix(counts)
    taxa, tree, kwargs = _get_phylogenetic_kwargs(kwargs, taxa)
    if kwargs:
        raise TypeError(f"faith_pd() got an unexpected keyword argument '{next(iter(kwargs))}'")
    if validate:
        validate_taxa_and_tree(taxa, tree)
    if _capi.load().b2d_device_count() <= 0:
        raise _capi.EngineError("no CUDA device visible: the B200 engine has no CPU fallback.")
    parent, length, names = flatten_tree(tree)
    # _nodes_by_counts truncates to int64 first (_phylogenetic.pyx:186); presence = count > 0
    indptr, cols, vals = to_csr(counts, "int64")
    keep = vals > 0
    if not keep.all():
        rows = np.repeat(np.arange(counts.shape[0]), np.diff(indptr))[keep]
        cols = cols[keep]
        indptr = np.zeros(counts.shape[0] + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=counts.shape[0]), out=indptr[1:])
    node_of_col = map_taxa_to_nodes(taxa, np.unique(cols), names)
    with _capi.Problem.tree(parent, length, indptr, node_of_col[cols], None, counts.shape[0],
                            device=device) as prob:
        prob.compute("faith_pd")
        vec = prob.sample_vector()
    return pd.Series(vec, index=ids)


def _phydiv_engine(counts, ids, validate, device, kwargs):
    """``alpha_diversity('phydiv', ..., rooted=None, weight=False)`` on the engine
    (/root/reference/skbio/diversity/_driver.py:192-204, alpha/_pd.py:190-239, 242-420)."""
    import pandas as pd

    taxa = kwargs.pop("taxa", None)
    counts, ids, taxa = ingest_table(counts, ids, taxa)
    if 0 in counts.shape:
        return pd.Series(np.zeros(len(ids)), index=ids)
    if validate:
        counts = validate_counts_matrix(counts)
    taxa, tree, kwargs = _get_phylogenetic_kwargs(kwargs, taxa)
    if isinstance(tree, FlatTree):
        default_rooted = int(np.count_nonzero(tree.parent == len(tree) - 1)) == 2
    else:
        default_rooted = len(_root_children(tree)) == 2     # TreeNode._is_rooted, _tree.py:385
    rooted = kwargs.pop("rooted", default_rooted)
    weight = kwargs.pop("weight", False)
    if kwargs:
        raise TypeError(f"_phydiv() got an unexpected keyword argument '{next(iter(kwargs))}'")
    if rooted is None:
        rooted = default_rooted
    if validate:
        validate_taxa_and_tree(taxa, tree, rooted=False)
    if isinstance(weight, bool):
        w = 1.0 if weight else 0.0
    else:
        w = float(weight)
        if not 0.0 <= w <= 1.0:
            raise ValueError("`weight` must be a boolean or a number within [0, 1].")
    if _capi.load().b2d_device_count() <= 0:
        raise _capi.EngineError("no CUDA device visible: the B200 engine has no CPU fallback.")
    parent, length, names = flatten_tree(tree)
    indptr, cols, vals = to_csr(counts, "int64")
    node_of_col = map_taxa_to_nodes(taxa, np.unique(cols), names)
    with _capi.Problem.tree(parent, length, indptr, node_of_col[cols], vals, counts.shape[0],
                            device=device) as prob:
        vec = prob.phydiv(bool(rooted), w)
    return pd.Series(vec, index=ids)


def _root_children(tree):
    r = getattr(tree, "root", None)
    node = r() if callable(r) else tree
    return node.children


def _reference_route(metric, counts, ids, validate, pairwise_func, taxa, kwargs):
    """Metrics that are not on the engine's path: run exactly what the reference runs."""
    if _is_sparse(counts):
        counts = counts.toarray()
    try:  # pragma: no cover - depends on the environment
        from skbio.diversity import beta_diversity as _skbio_beta_diversity  # type: ignore
    except Exception:
        _skbio_beta_diversity = None
    if _skbio_beta_diversity is not None:  # pragma: no cover
        kw = dict(kwargs)
        if taxa is not None:
            kw["taxa"] = taxa
        return _skbio_beta_diversity(metric, counts, ids, validate=validate,
                                     pairwise_func=pairwise_func, **kw)
    # scikit-bio absent: restate _driver.py:303-355 for SciPy / callable metrics
    if isinstance(metric, str) and metric in _qualitative_metrics:
        counts = counts > 0.0
    if metric in ("unweighted_unifrac", "weighted_unifrac"):
        raise _capi.EngineError(
            "a custom pairwise_func with a UniFrac metric needs scikit-bio's CPU "
            "implementation, which is not installed here.")
    if metric == "manhattan":
        metric = "cityblock"
    elif metric == "mahalanobis":
        nrow, ncol = counts.shape
        if nrow < ncol:
            raise ValueError(
                "Metric 'mahalanobis' requires more samples than features. "
                f"The input has {nrow} samples and {ncol} features.")
    elif callable(metric):
        if getattr(metric, "__name__", None) in _slow_beta_callables:
            warnings.warn(
                f"Passing `{metric.__name__}` as a callable invokes a much slower "
                "implementation. Pass the metric name as a string "
                f"(e.g., metric='{metric.__name__}') to use the optimized version.",
                stacklevel=3)
        if taxa is not None and "taxa" in signature(metric).parameters:
            kwargs["taxa"] = taxa
        metric = partial(metric, **kwargs)
        kwargs = {}
    elif metric not in _pdist_metrics:
        raise ValueError(
            f'"{metric}" is not an available beta diversity metric name. '
            "Refer to `get_beta_diversity_metrics` for a list of available metrics.")
    if pairwise_func is None:
        from scipy.spatial.distance import pdist

        pairwise_func = pdist
    distances = pairwise_func(counts, metric=metric, **kwargs)
    return make_distance_matrix(np.asarray(distances, dtype=np.float64), ids)


def install():
    """Rebind ``skbio.diversity.beta_diversity`` to this engine (optional monkey-patch)."""
    import skbio.diversity as _d  # type: ignore
    import skbio.diversity._driver as _drv  # type: ignore

    _d.beta_diversity = beta_diversity
    _drv.beta_diversity = beta_diversity
