"""ctypes binding of libskbb200.so (declared in include/skbio_b200.h).

Mirrors how the reference loads its optional accelerator:
``ctypes.CDLL("libskbb.so")`` cached on first use, a version probe, raw caller-owned NumPy
buffers passed as pointers (/root/reference/skbio/binaries/_util.py:28-110,
/root/reference/skbio/binaries/_distance.py:161-193).  There is NO CPU fallback here: when
the library or a GPU is missing, every compute entry point raises.
"""

from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_LIB = None
_LIB_LOCK = threading.Lock()

METRIC_IDS = {
    "unweighted_unifrac": 0,
    "weighted_unifrac": 1,
    "weighted_unifrac_normalized": 2,
    "braycurtis": 3,
    "jaccard": 4,
    "faith_pd": 5,
}

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)


# callbacks of b2d_problem_set_collectives (include/skbio_b200.h)
ALLGATHER_HOST_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)
ALLGATHERV_DEVICE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                        _i64p, _i64p, ctypes.c_int)


class EngineError(RuntimeError):
    """Raised when the native engine reports a failure (status code != 0)."""


def library_path():
    env = os.environ.get("SKBIO_B200_LIB")
    if env:
        return env
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libskbb200.so")


def _declare(lib):
    lib.b2d_api_version.restype = ctypes.c_int
    lib.b2d_device_count.restype = ctypes.c_int
    lib.b2d_last_error.restype = ctypes.c_char_p
    lib.b2d_host_alloc.restype = ctypes.c_void_p
    lib.b2d_host_alloc.argtypes = [ctypes.c_size_t]
    lib.b2d_host_free.restype = None
    lib.b2d_host_free.argtypes = [ctypes.c_void_p]
    lib.b2d_problem_create_tree.restype = ctypes.c_int
    lib.b2d_problem_create_tree.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
        _i32p, _f64p, _i64p, _i32p, _i64p, ctypes.c_int, ctypes.c_int]
    lib.b2d_problem_create_tree_columns.restype = ctypes.c_int
    lib.b2d_problem_create_tree_columns.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
        _i32p, _f64p, _i64p, _i32p, _i64p, _i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    lib.b2d_names_lookup.restype = ctypes.c_int
    lib.b2d_names_lookup.argtypes = [ctypes.c_char_p, _i64p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int64,
                                     ctypes.c_char_p, _i64p, ctypes.c_int64, ctypes.c_int, _i32p]
    lib.b2d_names_validate.restype = ctypes.c_int
    lib.b2d_names_validate.argtypes = [ctypes.c_char_p, _i64p, ctypes.POINTER(ctypes.c_uint8), _i32p,
                                       ctypes.c_int64, ctypes.c_char_p, _i64p, ctypes.c_int64, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int), _i64p]
    lib.b2d_problem_create_features.restype = ctypes.c_int
    lib.b2d_problem_create_features.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
        _i64p, _i32p, _f64p, ctypes.c_int, ctypes.c_int]
    lib.b2d_problem_select_tiles.restype = ctypes.c_int
    lib.b2d_problem_select_tiles.argtypes = [ctypes.c_void_p, _i32p, ctypes.c_int64]
    lib.b2d_problem_compute.restype = ctypes.c_int
    lib.b2d_problem_compute.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.b2d_problem_result_len.restype = ctypes.c_int64
    lib.b2d_problem_result_len.argtypes = [ctypes.c_void_p]
    lib.b2d_problem_fetch.restype = ctypes.c_int
    lib.b2d_problem_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.b2d_problem_fetch_compact.restype = ctypes.c_int
    lib.b2d_problem_fetch_compact.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _i64p, _i64p]
    lib.b2d_problem_sample_vector.restype = ctypes.c_int
    lib.b2d_problem_sample_vector.argtypes = [ctypes.c_void_p, _f64p]
    lib.b2d_problem_phydiv.restype = ctypes.c_int
    lib.b2d_problem_phydiv.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, _f64p]
    lib.b2d_problem_timings.restype = ctypes.c_int
    lib.b2d_problem_timings.argtypes = [ctypes.c_void_p, _f64p, ctypes.c_int]
    lib.b2d_problem_destroy.restype = ctypes.c_int
    lib.b2d_problem_destroy.argtypes = [ctypes.c_void_p]
    lib.b2d_beta_diversity_tree.restype = ctypes.c_int
    lib.b2d_beta_diversity_tree.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _i32p, _f64p, _i64p, _i32p,
        _i64p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.b2d_beta_diversity_features.restype = ctypes.c_int
    lib.b2d_beta_diversity_features.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _i64p, _i32p, _f64p,
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.b2d_set_option.restype = ctypes.c_int
    lib.b2d_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
    lib.b2d_microbench.restype = ctypes.c_int
    lib.b2d_microbench.argtypes = [ctypes.c_int, ctypes.c_int, _f64p]
    lib.b2d_permanova_sw.restype = ctypes.c_int
    lib.b2d_permanova_sw.argtypes = [ctypes.c_int, ctypes.c_int64, _f64p, ctypes.c_int64, _i32p,
                                     ctypes.c_int32, _f64p, _f64p]
    lib.b2d_newick_parse.restype = ctypes.c_int
    lib.b2d_newick_parse.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
    lib.b2d_flat_tree_nodes.restype = ctypes.c_int64
    lib.b2d_flat_tree_nodes.argtypes = [ctypes.c_void_p]
    lib.b2d_flat_tree_name_bytes.restype = ctypes.c_int64
    lib.b2d_flat_tree_name_bytes.argtypes = [ctypes.c_void_p]
    lib.b2d_flat_tree_get.restype = ctypes.c_int
    lib.b2d_flat_tree_get.argtypes = [ctypes.c_void_p, _i32p, _f64p, _i64p,
                                      ctypes.POINTER(ctypes.c_uint8), ctypes.c_char_p]
    lib.b2d_flat_tree_free.restype = None
    lib.b2d_flat_tree_free.argtypes = [ctypes.c_void_p]
    lib.b2d_shard_layout.restype = ctypes.c_int
    lib.b2d_shard_layout.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, _i64p, _i64p, _i64p]
    lib.b2d_problem_fetch_square_rows.restype = ctypes.c_int
    lib.b2d_problem_fetch_square_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, _f64p]
    lib.b2d_problem_center.restype = ctypes.c_int
    lib.b2d_problem_center.argtypes = [ctypes.c_void_p, _f64p, _f64p, _f64p]
    lib.b2d_center_condensed.restype = ctypes.c_int
    lib.b2d_center_condensed.argtypes = [ctypes.c_int, ctypes.c_int64, _f64p, _f64p, _f64p, _f64p]
    lib.b2d_debug_fetch_embedding.restype = ctypes.c_int
    lib.b2d_debug_fetch_embedding.argtypes = [ctypes.c_void_p, ctypes.c_int, _i64p]
    lib.b2d_problem_set_collectives.restype = ctypes.c_int
    lib.b2d_problem_set_collectives.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ALLGATHER_HOST_FN,
                                                ALLGATHERV_DEVICE_FN, ctypes.c_void_p]
    lib.b2d_build_range.restype = ctypes.c_int
    lib.b2d_build_range.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, _i64p, _i64p]
    lib.b2d_host_scan.restype = ctypes.c_int
    lib.b2d_host_scan.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, _f64p, _i64p, _i64p]
    lib.b2d_device_copy.restype = ctypes.c_int
    lib.b2d_device_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    lib.b2d_debug_tree_order.restype = ctypes.c_int
    lib.b2d_debug_tree_order.argtypes = [
        ctypes.c_int64, _i32p, _f64p, _i32p, ctypes.POINTER(ctypes.c_uint8), _f64p, _f64p,
        ctypes.POINTER(ctypes.c_int32)]


EXPORTED_SYMBOLS = [
    "b2d_api_version", "b2d_device_count", "b2d_last_error", "b2d_host_alloc", "b2d_host_free",
    "b2d_problem_create_tree", "b2d_problem_create_features", "b2d_problem_compute",
    "b2d_problem_result_len", "b2d_problem_fetch", "b2d_problem_fetch_compact",
    "b2d_problem_sample_vector", "b2d_problem_timings", "b2d_problem_destroy", "b2d_debug_tree_paths",
    "b2d_beta_diversity_tree", "b2d_beta_diversity_features", "b2d_set_option",
    "b2d_microbench", "b2d_debug_tree_order", "b2d_shard_layout", "b2d_problem_select_tiles",
    "b2d_newick_parse", "b2d_flat_tree_nodes", "b2d_flat_tree_name_bytes", "b2d_flat_tree_get",
    "b2d_flat_tree_free", "b2d_permanova_sw", "b2d_problem_phydiv", "b2d_debug_fetch_embedding",
    "b2d_problem_fetch_square_rows", "b2d_problem_center", "b2d_center_condensed",
    "b2d_problem_create_tree_columns", "b2d_names_lookup", "b2d_names_validate",
    "b2d_problem_set_collectives", "b2d_build_range", "b2d_device_copy", "b2d_host_scan",
]


def load():
    """Load (once) and return the native library; raises EngineError if it is missing."""
    global _LIB
    with _LIB_LOCK:
        if _LIB is None:
            path = library_path()
            if not os.path.exists(path):
                raise EngineError(
                    f"native engine not built: {path} is missing. Run "
                    "`python -c 'import __graft_entry__ as g; g.build()'` or "
                    "`make -C scikit-bio_b200/csrc`. There is no CPU fallback.")
            lib = ctypes.CDLL(path)
            _declare(lib)
            _LIB = lib
    return _LIB


def available():
    """True when the library loads and at least one CUDA device is visible (mirrors
    skbio.binaries.available)."""
    try:
        return load().b2d_device_count() > 0
    except Exception:
        return False


def _check(rc):
    if rc != 0:
        msg = load().b2d_last_error().decode("utf-8", "replace")
        raise EngineError(f"libskbb200 error {rc}: {msg}")


def _ptr(arr, typ):
    if arr is None:
        return None
    return arr.ctypes.data_as(typ)


def set_option(name, value):
    _check(load().b2d_set_option(name.encode(), int(value)))


def release_cache():
    """Give every device buffer the library has parked back to the CUDA driver.

    Freed buffers of a finished problem are parked per device (up to ``cache_limit_bytes``,
    4 GiB by default) because cudaMalloc / cudaFree of large blocks cost 0.1 - 0.7 s; call this
    before handing the GPU to another library in the same process."""
    set_option("release_cache", 1)


def set_cache_limit(nbytes):
    """Upper bound (bytes per device) on parked device memory; 0 parks nothing.  Loops that
    create and destroy many large problems (benchmarks, block-wise runs) may raise it so that
    the embedding buffer itself is reused."""
    set_option("cache_limit_bytes", int(nbytes))


def microbench(kind, device=0):
    out = ctypes.c_double(0.0)
    _check(load().b2d_microbench(int(device), int(kind), ctypes.byref(out)))
    return out.value


def permanova_sw(condensed, groupings, inv_group_size, device=0):
    """s_W for every row of ``groupings`` ([k][n] int32) on a condensed distance vector."""
    lib = load()
    condensed = np.ascontiguousarray(condensed, dtype=np.float64)
    groupings = np.ascontiguousarray(groupings, dtype=np.int32)
    inv = np.ascontiguousarray(inv_group_size, dtype=np.float64)
    k, n = groupings.shape
    out = np.empty(k, dtype=np.float64)
    _check(lib.b2d_permanova_sw(int(device), n, _ptr(condensed, _f64p), k, _ptr(groupings, _i32p),
                                len(inv), _ptr(inv, _f64p), _ptr(out, _f64p)))
    return out


def center_condensed(condensed, device=0, want_matrix=True):
    """(F [n x n] or None, row means of E [n], global mean of E) for a condensed distance vector."""
    lib = load()
    condensed = np.ascontiguousarray(condensed, dtype=np.float64)
    ln = len(condensed)
    n = int(round((1.0 + np.sqrt(1.0 + 8.0 * ln)) / 2.0))
    if n * (n - 1) // 2 != ln:
        raise ValueError("not a condensed vector: its length is not n(n-1)/2")
    rm = np.empty(n, dtype=np.float64)
    gm = ctypes.c_double(0.0)
    f = np.empty((n, n), dtype=np.float64) if want_matrix else None
    _check(lib.b2d_center_condensed(int(device), n, _ptr(condensed, _f64p), _ptr(rm, _f64p), ctypes.byref(gm),
                                    _ptr(f, _f64p)))
    return f, rm, gm.value


class NameBlob:
    """Node names as one UTF-8 byte buffer with offsets (no per-node Python strings)."""

    __slots__ = ("buf", "off", "has")

    def __init__(self, buf, off, has):
        self.buf = buf                                   # bytes
        self.off = np.ascontiguousarray(off, dtype=np.int64)   # [m + 1]
        self.has = np.ascontiguousarray(has, dtype=np.uint8)   # [m]

    def decode(self):
        m = len(self.has)
        text = self.buf.decode("utf-8")
        offl = self.off.tolist()
        if len(text) != len(self.buf):   # multi-byte characters: slice the bytes
            names = [self.buf[offl[k]:offl[k + 1]].decode("utf-8") for k in range(m)]
        else:
            names = [text[offl[k]:offl[k + 1]] for k in range(m)]
        out = np.empty(m, dtype=object)
        out[:] = names
        out[self.has == 0] = None
        return out


def query_blob(names):
    """(bytes, offsets int64[k + 1]) of a list of strings."""
    try:
        joined = "\0".join(names)
    except TypeError:                    # non-string taxa (numbers, ...): their str() as the reference does
        names = [x if isinstance(x, str) else str(x) for x in names]
        joined = "\0".join(names)
    buf = joined.encode("utf-8") + b"\0"
    off = np.zeros(len(names) + 1, dtype=np.int64)
    # entry k is buf[off[k] : off[k + 1] - 1]: the offsets are the positions after the separators
    ends = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 0)
    if len(ends) == len(names):
        off[1:] = ends + 1
    elif len(names):                     # a name contains a NUL byte itself: count bytes name by name
        np.cumsum([len(x.encode("utf-8")) + 1 for x in names], out=off[1:])
    # the separators are part of the buffer but of no entry
    return buf, off


def names_lookup(blob, queries, qblob=None):
    """Node id (last node carrying the name wins) or -1 for every query string - host only."""
    qb, qo = qblob if qblob is not None else query_blob(queries)
    out = np.empty(len(queries), dtype=np.int32)
    _check(load().b2d_names_lookup(blob.buf, _ptr(blob.off, _i64p),
                                   blob.has.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), len(blob.has),
                                   qb, _ptr(qo, _i64p), len(queries), 1, _ptr(out, _i32p)))
    return out


def names_validate(blob, parent, queries, qblob=None):
    """(a tip name occurs twice, number of distinct queries that are not tip names) - host only."""
    qb, qo = qblob if qblob is not None else query_blob(queries)
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    dup, missing = ctypes.c_int(0), ctypes.c_int64(0)
    _check(load().b2d_names_validate(blob.buf, _ptr(blob.off, _i64p),
                                     blob.has.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                     _ptr(parent, _i32p), len(blob.has), qb, _ptr(qo, _i64p), len(queries), 1,
                                     ctypes.byref(dup), ctypes.byref(missing)))
    return bool(dup.value), int(missing.value)


def newick_to_flat(text):
    """(parent int32[m], length float64[m] with NaN for 'no length', NameBlob) straight from
    Newick text, in the reference's node id order - no node objects (host only)."""
    lib = load()
    raw = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    h = ctypes.c_void_p()
    _check(lib.b2d_newick_parse(raw, len(raw), ctypes.byref(h)))
    try:
        m = int(lib.b2d_flat_tree_nodes(h))
        nb = int(lib.b2d_flat_tree_name_bytes(h))
        parent = np.empty(m, dtype=np.int32)
        length = np.empty(m, dtype=np.float64)
        off = np.empty(m + 1, dtype=np.int64)
        has = np.empty(m, dtype=np.uint8)
        buf = ctypes.create_string_buffer(max(nb, 1))
        _check(lib.b2d_flat_tree_get(h, _ptr(parent, _i32p), _ptr(length, _f64p), _ptr(off, _i64p),
                                     has.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), buf))
    finally:
        lib.b2d_flat_tree_free(h)
    return parent, length, NameBlob(buf.raw[:nb], off, has)


def build_range(n_samples, rank, world):
    """Blocks [first, last) of 128 samples whose per-block products rank `rank` builds - host only."""
    a, b = ctypes.c_int64(0), ctypes.c_int64(0)
    _check(load().b2d_build_range(int(n_samples), int(rank), int(world), ctypes.byref(a), ctypes.byref(b)))
    return int(a.value), int(b.value)


_SCAN_DTYPES = {np.dtype(np.int64): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.float32): 3}


def host_scan(values):
    """(minimum, number of non-zero, number of positive) elements of a 1-D array - one multi-threaded
    native pass for large contiguous int64 / float64 / int32 / float32 arrays, NumPy otherwise.  Host only."""
    values = np.asarray(values)
    code = _SCAN_DTYPES.get(values.dtype)
    if code is None or values.ndim != 1 or not values.flags.c_contiguous or values.size < (1 << 16):
        if values.size == 0:
            return 0.0, 0, 0
        return float(values.min()), int(np.count_nonzero(values)), int(np.count_nonzero(values > 0))
    mn = ctypes.c_double(0.0)
    nz, pos = ctypes.c_int64(0), ctypes.c_int64(0)
    _check(load().b2d_host_scan(values.ctypes.data, values.size, code, ctypes.byref(mn), ctypes.byref(nz),
                                ctypes.byref(pos)))
    return mn.value, int(nz.value), int(pos.value)


def device_copy(dst, src, nbytes):
    """Blocking device-to-device copy between raw pointers (same or peer device)."""
    _check(load().b2d_device_copy(ctypes.c_void_p(int(dst)), ctypes.c_void_p(int(src)), int(nbytes)))


def shard_layout(n_samples, shard, n_shards):
    """(offsets[k, 2] = {first condensed index, count} per owned row stripe, total) - host only."""
    lib = load()
    nb = (int(n_samples) + 127) // 128
    offs = np.zeros(2 * max(nb, 1), dtype=np.int64)
    cnt, tot = ctypes.c_int64(0), ctypes.c_int64(0)
    _check(lib.b2d_shard_layout(int(n_samples), int(shard), int(n_shards), _ptr(offs, _i64p),
                                ctypes.byref(cnt), ctypes.byref(tot)))
    return offs[: 2 * cnt.value].reshape(-1, 2).copy(), int(tot.value)


def debug_tree_order(parent, length):
    """Engine node order of a flat tree (host-only test hook)."""
    lib = load()
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    length = np.ascontiguousarray(length, dtype=np.float64)
    m = len(parent)
    mpad = (m + 127) // 128 * 128
    q = np.empty(m, dtype=np.int32)
    flags = np.empty(mpad, dtype=np.uint8)
    lq = np.empty(mpad, dtype=np.float64)
    td = np.empty(mpad, dtype=np.float64)
    depth = ctypes.c_int32(0)
    _check(lib.b2d_debug_tree_order(m, _ptr(parent, _i32p), _ptr(length, _f64p), _ptr(q, _i32p),
                                    flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                    _ptr(lq, _f64p), _ptr(td, _f64p), ctypes.byref(depth)))
    return {"q_of_id": q, "flags": flags, "len_q": lq, "tipdist_q": td, "max_depth": depth.value}


def debug_tree_paths(parent, length):
    """Per-position parent / subtree size / double-double root distance (host-only test hook)."""
    lib = load()
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    length = np.ascontiguousarray(length, dtype=np.float64)
    m = len(parent)
    mpad = (m + 127) // 128 * 128
    pq = np.empty(mpad, dtype=np.int32)
    sz = np.empty(mpad, dtype=np.int32)
    hi = np.empty(mpad, dtype=np.float64)
    lo = np.empty(mpad, dtype=np.float64)
    tq = np.empty(mpad, dtype=np.int32)
    hi_int = np.empty(mpad, dtype=np.float64)
    lo_int = np.empty(mpad, dtype=np.float64)
    lib.b2d_debug_tree_paths.restype = ctypes.c_int
    lib.b2d_debug_tree_paths.argtypes = [ctypes.c_int64, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p, _i32p,
                                         _f64p, _f64p]
    _check(lib.b2d_debug_tree_paths(m, _ptr(parent, _i32p), _ptr(length, _f64p), _ptr(pq, _i32p),
                                    _ptr(sz, _i32p), _ptr(hi, _f64p), _ptr(lo, _f64p), _ptr(tq, _i32p),
                                    _ptr(hi_int, _f64p), _ptr(lo_int, _f64p)))
    return {"parent_q": pq, "size_q": sz, "rd_hi": hi, "rd_lo": lo, "tips_q": tq, "rd_hi_int": hi_int,
            "rd_lo_int": lo_int}


class PinnedBuffer:
    """Vector (fp64 by default) in page-locked host memory owned by the native library."""

    def __init__(self, n, dtype=np.float64):
        self._lib = load()
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.nbytes = max(self.n, 1) * self.dtype.itemsize
        self.ptr = self._lib.b2d_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("b2d_host_alloc failed")
        self.array = self.new_array()

    def new_array(self):
        """A fresh ndarray over the buffer (the buffer itself keeps no reference to it when the
        caller clears `.array`, so its lifetime can be tracked with a weakref)."""
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        return np.frombuffer(buf, dtype=self.dtype, count=self.n)

    @classmethod
    def like(cls, a):
        """Page-locked copy of a host array (so that its upload is a plain DMA)."""
        a = np.ascontiguousarray(a)
        b = cls(a.size, a.dtype)
        b.array[:] = a.ravel()
        return b

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.b2d_host_free(self.ptr)
            self.ptr = None
            self.n = 0

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


class Problem:
    """Inputs of one beta-diversity call resident on one GPU (one shard of the triangle)."""

    def __init__(self, handle, n):
        self._h = handle
        self.n = n
        self._keep = None

    @classmethod
    def tree(cls, parent, length, indptr, node_ids, values, n_samples, device=0, shard=0,
             n_shards=1, node_of_col=None):
        """`node_ids` are node ids, or - with `node_of_col` - table columns that the device maps
        through that table (column -> node id)."""
        lib = load()
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        length = np.ascontiguousarray(length, dtype=np.float64)
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        node_ids = np.ascontiguousarray(node_ids, dtype=np.int32)
        if values is not None:
            values = np.ascontiguousarray(values, dtype=np.int64)
        h = ctypes.c_void_p()
        if node_of_col is not None:
            node_of_col = np.ascontiguousarray(node_of_col, dtype=np.int32)
            _check(lib.b2d_problem_create_tree_columns(
                ctypes.byref(h), int(device), int(n_samples), int(len(parent)),
                _ptr(parent, _i32p), _ptr(length, _f64p), _ptr(indptr, _i64p),
                _ptr(node_ids, _i32p), _ptr(values, _i64p), _ptr(node_of_col, _i32p), len(node_of_col),
                int(shard), int(n_shards)))
        else:
            _check(lib.b2d_problem_create_tree(
                ctypes.byref(h), int(device), int(n_samples), int(len(parent)),
                _ptr(parent, _i32p), _ptr(length, _f64p), _ptr(indptr, _i64p),
                _ptr(node_ids, _i32p), _ptr(values, _i64p), int(shard), int(n_shards)))
        return cls(h, int(n_samples))

    @classmethod
    def features(cls, indptr, feature_ids, values, n_samples, n_features, device=0, shard=0,
                 n_shards=1):
        lib = load()
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        feature_ids = np.ascontiguousarray(feature_ids, dtype=np.int32)
        if values is not None:
            values = np.ascontiguousarray(values, dtype=np.float64)
        h = ctypes.c_void_p()
        _check(lib.b2d_problem_create_features(
            ctypes.byref(h), int(device), int(n_samples), int(n_features),
            _ptr(indptr, _i64p), _ptr(feature_ids, _i32p), _ptr(values, _f64p),
            int(shard), int(n_shards)))
        return cls(h, int(n_samples))

    def select_tiles(self, tiles):
        """Restrict to 128x128 tiles given as an (k, 2) array of (row stripe, column stripe)."""
        tiles = np.ascontiguousarray(tiles, dtype=np.int32).reshape(-1, 2)
        _check(load().b2d_problem_select_tiles(self._h, _ptr(tiles, _i32p), len(tiles)))

    def set_collectives(self, rank, world, coll):
        """Distributed build of the per-block products (b2d_problem_set_collectives).  `coll` offers
        ``allgather_host(mine: bytes) -> bytes`` (the ranks' blobs back to back) and
        ``allgatherv_device(bases, offsets, counts)`` (lists: one device pointer per buffer, byte
        offsets / counts per buffer and rank; in place).  See skbio_b200.distributed."""
        if coll is None or world <= 1:
            _check(load().b2d_problem_set_collectives(self._h, 0, 1, ALLGATHER_HOST_FN(0), ALLGATHERV_DEVICE_FN(0), None))
            self._coll = None
            return
        errors = []

        def host_cb(user, mine, out, nbytes):
            try:
                blob = coll.allgather_host(ctypes.string_at(mine, nbytes))
                if len(blob) != nbytes * world:
                    raise ValueError("allgather_host returned %d bytes, expected %d" % (len(blob), nbytes * world))
                ctypes.memmove(out, blob, len(blob))
                return 0
            except BaseException as e:   # must not propagate through the C frames
                errors.append(e)
                return 1

        def dev_cb(user, nbuf, bases, offsets, counts, w):
            try:
                coll.allgatherv_device([int(bases[k] or 0) for k in range(nbuf)],
                                       [[int(offsets[k * w + r]) for r in range(w)] for k in range(nbuf)],
                                       [[int(counts[k * w + r]) for r in range(w)] for k in range(nbuf)])
                return 0
            except BaseException as e:
                errors.append(e)
                return 1

        self._coll = (ALLGATHER_HOST_FN(host_cb), ALLGATHERV_DEVICE_FN(dev_cb), coll, errors)
        _check(load().b2d_problem_set_collectives(self._h, int(rank), int(world), self._coll[0], self._coll[1], None))

    def compute(self, metric):
        mid = METRIC_IDS[metric] if isinstance(metric, str) else int(metric)
        coll = getattr(self, "_coll", None)
        if coll is not None:
            del coll[3][:]
            rc = load().b2d_problem_compute(self._h, mid)
            if rc != 0 and coll[3]:
                raise EngineError("collective callback failed: %r" % (coll[3][0],)) from coll[3][0]
            _check(rc)
            return
        _check(load().b2d_problem_compute(self._h, mid))

    def result_len(self):
        return int(load().b2d_problem_result_len(self._h))

    def fetch(self, out):
        """Write the owned slices into the full condensed vector ``out`` (float64)."""
        assert out.dtype == np.float64 and out.flags.c_contiguous
        assert out.size == self.n * (self.n - 1) // 2
        if out.size:
            _check(load().b2d_problem_fetch(self._h, out.ctypes.data))
        return out

    def fetch_compact(self, out=None):
        lib = load()
        m = self.result_len()
        if out is None:
            out = np.empty(m, dtype=np.float64)
        nb = (self.n + 127) // 128
        offs = np.zeros(2 * max(nb, 1), dtype=np.int64)
        cnt = ctypes.c_int64(0)
        _check(lib.b2d_problem_fetch_compact(self._h, out.ctypes.data, _ptr(offs, _i64p),
                                             ctypes.byref(cnt)))
        return out, offs[: 2 * cnt.value].reshape(-1, 2)

    def fetch_square_rows(self, r0, r1):
        """Rows [r0, r1) of the square distance matrix, gathered on the device (single shard)."""
        out = np.empty((int(r1) - int(r0), self.n), dtype=np.float64)
        if out.size:
            _check(load().b2d_problem_fetch_square_rows(self._h, int(r0), int(r1), _ptr(out, _f64p)))
        return out

    def center(self, want_matrix=True):
        """Gower centering of the device-resident result: (F or None, row means, global mean)."""
        rm = np.empty(self.n, dtype=np.float64)
        gm = ctypes.c_double(0.0)
        f = np.empty((self.n, self.n), dtype=np.float64) if want_matrix else None
        _check(load().b2d_problem_center(self._h, _ptr(rm, _f64p), ctypes.byref(gm), _ptr(f, _f64p)))
        return f, rm, gm.value

    def sample_vector(self):
        out = np.empty(self.n, dtype=np.float64)
        if self.n:
            _check(load().b2d_problem_sample_vector(self._h, _ptr(out, _f64p)))
        return out

    def phydiv(self, rooted, weight):
        out = np.empty(self.n, dtype=np.float64)
        if self.n:
            _check(load().b2d_problem_phydiv(self._h, 1 if rooted else 0, float(weight), _ptr(out, _f64p)))
        return out

    def fetch_embedding(self, variant=0, n_nodes=None):
        """Test hook: the device-built embedding as (n_samples x n_nodes) int64 in reference id
        order (variant 0 / 1: subtree counts, sequential / decomposed walk; 2: presence)."""
        out = np.zeros((self.n, int(n_nodes)), dtype=np.int64)
        if out.size:
            _check(load().b2d_debug_fetch_embedding(self._h, int(variant), _ptr(out, _i64p)))
        return out

    def timings(self):
        t = np.zeros(21, dtype=np.float64)
        _check(load().b2d_problem_timings(self._h, _ptr(t, _f64p), 21))
        return {"embed_ms": t[0], "pair_ms": t[1], "compute_ms": t[2], "fetch_ms": t[3],
                "upload_ms": t[4], "pair_launches": int(t[5]), "nodes_resident": int(t[6]),
                "passes": int(t[7]), "visits_executed": float(t[8]), "visits_dense": float(t[9]),
                "skip_prep_ms": t[10], "kernels": int(t[11]), "tips_ms": t[12],
                "tips_flagged": int(t[13]), "sparse_rows_max_tips": int(t[14]),
                "sparse_row_entries": int(t[15]), "sparse_row_matches": float(t[16]),
                "isect_generation": int(t[17]), "exchange_ms": t[18], "exchange_bytes": float(t[19]),
                "csr_upload_bytes": float(t[20])}

    def destroy(self):
        if self._h is not None:
            load().b2d_problem_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass
