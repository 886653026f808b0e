"""Counts-table ingestion, validation and CSR conversion (host side).

Mirrors, for this path only:

* ``_ingest_table``            /root/reference/skbio/table/_tabular.py:114-219
* ``_validate_counts_matrix``  /root/reference/skbio/diversity/_util.py:94-120 (and :15-64)
* ``_qualify_counts``          /root/reference/skbio/diversity/_util.py:184-185
* the int64 cast of ``_nodes_by_counts``  /root/reference/skbio/diversity/_phylogenetic.pyx:186

The reference densifies every input; here a ``scipy.sparse`` matrix (or a BIOM-like object
exposing ``matrix_data``) stays sparse all the way to the device.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np


def _is_sparse(x):
    return hasattr(x, "tocsr") and hasattr(x, "nnz")


def ingest_table(table, sample_ids=None, feature_ids=None, expand=True):
    """Return ``(data, sample_ids, feature_ids)``; ``data`` is a 2-D ndarray or a CSR matrix."""
    data, samples, features = None, None, None
    if _is_sparse(table):
        data = table.tocsr()
    elif isinstance(table, Sequence) and not isinstance(table, (str, bytes)):
        data = np.asarray(table)
    elif isinstance(table, np.ndarray):
        data = table
    else:
        try:
            import pandas as pd
        except Exception:  # pragma: no cover
            pd = None
        if pd is not None and isinstance(table, pd.DataFrame):
            data = table.to_numpy()
            samples = table.index
            features = table.columns
        elif hasattr(table, "generated_by") and hasattr(table, "matrix_data"):
            # BIOM table: features x samples sparse matrix -> samples x features CSR
            data = table.matrix_data.T.tocsr()
            samples = table.ids()
            features = table.ids("observation")
        elif hasattr(table, "X") and hasattr(table, "obs") and hasattr(table, "var"):
            x = table.X
            data = x.tocsr() if _is_sparse(x) else np.asarray(x)
            samples = table.obs.index
            features = table.var.index
        elif hasattr(table, "schema") and hasattr(table, "to_numpy"):
            data = table.to_numpy()
            features = table.schema
    if data is None:
        data = np.asarray(table)
    if not _is_sparse(data):
        if data.ndim == 0:
            raise TypeError(f"'{table.__class__.__name__}' is not a supported table format.")
        if data.ndim == 1 and expand:
            data = data.reshape(1, -1)
        if data.ndim < 2:
            raise ValueError("Input table has less than 2 dimensions.")
    lenerr = "Input table has {0} {1}s whereas {2} {1} IDs were provided."
    if sample_ids is None:
        sample_ids = samples
    elif len(sample_ids) != data.shape[0]:
        raise ValueError(lenerr.format(data.shape[0], "sample", len(sample_ids)))
    if feature_ids is None:
        feature_ids = features
    elif len(feature_ids) != data.shape[1]:
        raise ValueError(lenerr.format(data.shape[1], "feature", len(feature_ids)))
    return data, sample_ids, feature_ids


def _stored_values_scan(data):
    """(min, non-zero, positive) of a sparse table's stored values: one native multi-threaded pass when the
    library is there (host code only, no GPU needed), NumPy otherwise."""
    try:
        from . import _capi
        return _capi.host_scan(data)
    except Exception:
        data = np.asarray(data)
        if data.size == 0:
            return 0.0, 0, 0
        return float(data.min()), int(np.count_nonzero(data)), int(np.count_nonzero(data > 0))


def validate_counts_matrix(counts):
    """dtype / negativity / dimensionality checks with the reference's messages."""
    if _is_sparse(counts):
        dtype = counts.dtype
        if not (np.issubdtype(dtype, np.floating) or np.issubdtype(dtype, np.integer)
                or dtype == np.dtype("bool")):
            raise ValueError("Counts must be integers or floating-point numbers.")
        if counts.nnz and _stored_values_scan(counts.data)[0] < 0:
            raise ValueError("Counts cannot contain negative values.")
        return counts
    counts = np.asarray(counts)
    dtype = counts.dtype
    if np.issubdtype(dtype, np.floating):
        pass
    elif not np.issubdtype(dtype, np.integer) and dtype is not np.dtype("bool"):
        raise ValueError("Counts must be integers or floating-point numbers.")
    if counts.size > 0 and counts.min() < 0:
        raise ValueError("Counts cannot contain negative values.")
    counts = np.atleast_2d(counts)
    if counts.ndim > 2:
        raise ValueError(
            f"`counts` has {counts.ndim} dimensions whereas up to 2 dimensions are allowed.")
    return counts


def to_csr(counts, mode):
    """CSR triple ``(indptr int64, cols int32, values)`` of the non-zero entries.

    mode:
      'presence'  entries with ``counts > 0.0``; values None            (qualitative metrics)
      'int64'     ``counts.astype(int64)`` (truncation), non-zero kept  (weighted UniFrac)
      'float64'   ``counts`` as double, non-zero kept                   (Bray-Curtis)
    """
    if _is_sparse(counts):
        csr = counts.tocsr()
        data = csr.data
        # the common case - every stored entry counts - passes the caller's arrays on untouched
        if mode == "presence":
            nkeep = _stored_values_scan(data)[2]
        elif mode == "int64":
            data = data.astype(np.int64, copy=False)
            nkeep = _stored_values_scan(data)[1]
        else:
            data = data.astype(np.float64, copy=False)
            nkeep = _stored_values_scan(data)[1]
        indptr = csr.indptr.astype(np.int64)
        cols = csr.indices
        if nkeep != len(data):
            keep = (csr.data > 0.0) if mode == "presence" else (data != 0)
            rows = np.repeat(np.arange(csr.shape[0]), np.diff(indptr))[keep]
            cols = cols[keep]
            data = data[keep]
            indptr = np.zeros(csr.shape[0] + 1, dtype=np.int64)
            np.cumsum(np.bincount(rows, minlength=csr.shape[0]), out=indptr[1:])
        vals = None if mode == "presence" else np.ascontiguousarray(data)
        return indptr, np.ascontiguousarray(cols, dtype=np.int32), vals
    counts = np.atleast_2d(np.asarray(counts))
    if mode == "presence":
        work = counts > 0.0
    elif mode == "int64":
        work = counts.astype(np.int64, copy=False)
    else:
        work = counts.astype(np.float64, copy=False)
    rows, cols = np.nonzero(work)
    n = counts.shape[0]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    vals = None if mode == "presence" else np.ascontiguousarray(work[rows, cols])
    return indptr, cols.astype(np.int32), vals
