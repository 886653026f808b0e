"""ctypes wrapper of the C oracle (oracle_c.c).  TEST INFRASTRUCTURE ONLY - see oracle_np.py.

Inputs here are numeric (the taxon -> node map is given as ``node_of_col``) so mid-size
random cases can be checked without Python-level name handling.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads():
    return int(lib().orc_max_threads())


def _p(a, t):
    return a.ctypes.data_as(t)


def counts_by_node(dense_counts, node_of_col, parent):
    """(n_samples x n_nodes) int64, contiguous."""
    L = lib()
    counts = np.ascontiguousarray(dense_counts, dtype=np.int64)
    n, t = counts.shape
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    node_of_col = np.ascontiguousarray(node_of_col, dtype=np.int32)
    m = len(parent)
    a = np.zeros((m, n), dtype=np.int64)
    L.orc_nodes_by_counts(ctypes.c_int64(n), ctypes.c_int64(t), ctypes.c_int64(m), _p(counts, _i64p),
                          _p(node_of_col, _i32p), _p(parent, _i32p), _p(a, _i64p))
    return np.ascontiguousarray(a.T)


def tip_mask(parent):
    parent = np.asarray(parent)
    has_child = np.zeros(len(parent), dtype=bool)
    has_child[parent[parent >= 0]] = True
    return (~has_child).astype(np.uint8)


def tip_distances(length, parent):
    d = np.array(length, dtype=np.float64, copy=True)
    for i in range(len(d) - 1, -1, -1):
        if parent[i] >= 0:
            d[i] += d[parent[i]]
    return d * tip_mask(parent)


def unweighted(at, bl):
    n, m = at.shape
    out = np.empty(n * (n - 1) // 2, dtype=np.float64)
    lib().orc_unweighted(ctypes.c_int64(n), ctypes.c_int64(m), _p(at, _i64p), _p(bl, _f64p), _p(out, _f64p))
    return out


def weighted(at, bl, parent, normalized):
    n, m = at.shape
    out = np.empty(n * (n - 1) // 2, dtype=np.float64)
    tm = tip_mask(parent)
    d = np.ascontiguousarray(tip_distances(bl, parent)) if normalized else np.zeros(1)
    lib().orc_weighted(ctypes.c_int64(n), ctypes.c_int64(m), _p(at, _i64p), _p(bl, _f64p), _p(tm, _u8p),
                       ctypes.c_int(1 if normalized else 0), _p(d, _f64p), _p(out, _f64p))
    return out


def faith_pd(at, bl):
    n, m = at.shape
    out = np.empty(n, dtype=np.float64)
    lib().orc_faith_pd(ctypes.c_int64(n), ctypes.c_int64(m), _p(at, _i64p), _p(bl, _f64p), _p(out, _f64p))
    return out


def braycurtis(dense_counts):
    x = np.ascontiguousarray(dense_counts, dtype=np.float64)
    n, f = x.shape
    out = np.empty(n * (n - 1) // 2, dtype=np.float64)
    lib().orc_braycurtis(ctypes.c_int64(n), ctypes.c_int64(f), _p(x, _f64p), _p(out, _f64p))
    return out


def jaccard(dense_counts):
    x = np.ascontiguousarray(dense_counts, dtype=np.float64)
    n, f = x.shape
    out = np.empty(n * (n - 1) // 2, dtype=np.float64)
    lib().orc_jaccard(ctypes.c_int64(n), ctypes.c_int64(f), _p(x, _f64p), _p(out, _f64p))
    return out


def unifrac_all(dense_counts, node_of_col, parent, length):
    """All five metrics + Faith PD for a dense integer table (columns mapped by node_of_col)."""
    bl = np.ascontiguousarray(length, dtype=np.float64)
    at = counts_by_node(dense_counts, node_of_col, parent)
    at_presence = counts_by_node((np.asarray(dense_counts) > 0).astype(np.int64), node_of_col, parent)
    return {
        "unweighted_unifrac": unweighted(at_presence, bl),
        "weighted_unifrac": weighted(at, bl, parent, False),
        "weighted_unifrac_normalized": weighted(at, bl, parent, True),
        "faith_pd": faith_pd(at_presence, bl),
        "braycurtis": braycurtis(dense_counts),
        "jaccard": jaccard(dense_counts),
    }
