"""Streaming writers for the distance matrix ("next" row f4 of SURVEY.md section 8).

The reference writes a ``DistanceMatrix`` from its in-memory n x n array
(``_matrix_to_lsmat``, /root/reference/skbio/io/format/lsmat.py:266-279; ``binary_dm``'s
``_skbio_mat_to_h5py_mat``, /root/reference/skbio/io/format/binary_dm.py:157-184).  At 100,000
samples that array is 80 GB.  Here the file is produced row block by row block from

* a computed engine problem (`_capi.Problem`): the rows are gathered ON THE DEVICE from the
  resident condensed result (``b2d_problem_fetch_square_rows``) - neither the square nor the
  condensed matrix ever exists on the host;
* a condensed vector or any object with ``condensed_form()`` / ``ids``.

File contents are the reference's: lsmat = a header line of delimiter-prefixed ids, then one line
per object (id, then the row formatted by ``numpy`` ``str``); binary_dm = HDF5 datasets ``format``,
``version``, ``order`` (UTF-8 ids) and ``matrix`` (n x n float64) - needs ``h5py``.
"""

from __future__ import annotations

import numpy as np

from . import _capi


def _square_rows_from_condensed(vec, n, r0, r1):
    out = np.zeros((r1 - r0, n), dtype=np.float64)
    for i in range(r0, r1):
        row = out[i - r0]
        start = i * n - (i * (i + 1)) // 2          # index of (i, i + 1)
        row[i + 1:] = vec[start:start + n - 1 - i]
        if i:
            j = np.arange(i, dtype=np.int64)
            row[:i] = vec[j * n - ((j + 2) * (j + 1)) // 2 + i]
    return out


class _RowSource:
    """Row blocks of the square matrix from whichever form the distances are in."""

    def __init__(self, source, ids):
        self.problem = source if isinstance(source, _capi.Problem) else None
        if self.problem is not None:
            self.n = self.problem.n
            self.vec = None
        else:
            vec = source.condensed_form() if hasattr(source, "condensed_form") else source
            self.vec = np.ascontiguousarray(vec, dtype=np.float64)
            if self.vec.ndim != 1:
                raise ValueError("expected a condensed vector, a DistanceMatrix or an engine problem")
            ln = len(self.vec)
            self.n = int(round((1.0 + np.sqrt(1.0 + 8.0 * ln)) / 2.0)) if ln else 1
            if self.n * (self.n - 1) // 2 != ln:
                raise ValueError("not a condensed vector: its length is not n(n-1)/2")
            if ids is None and hasattr(source, "ids"):
                ids = source.ids
        if ids is None:
            ids = [str(i) for i in range(self.n)]
        self.ids = list(ids)
        if len(self.ids) != self.n:
            raise ValueError(f"The number of IDs ({len(self.ids)}) must match the number of "
                             f"rows/columns in the data ({self.n}).")

    def rows(self, r0, r1):
        if self.problem is not None:
            return self.problem.fetch_square_rows(r0, r1)
        return _square_rows_from_condensed(self.vec, self.n, r0, r1)

    def blocks(self, block_rows):
        for r0 in range(0, self.n, block_rows):
            r1 = min(self.n, r0 + block_rows)
            yield r0, r1, self.rows(r0, r1)


def write_lsmat(source, file, ids=None, delimiter="\t", block_rows=256):
    """Write the distances as a labelled square matrix (the reference's ``lsmat`` format), one
    block of rows at a time.  `file` is a path or a text file object."""
    src = _RowSource(source, ids)
    delimiter = "%s" % delimiter
    own = isinstance(file, (str, bytes)) or hasattr(file, "__fspath__")
    fh = open(file, "w") if own else file
    try:
        fh.write(delimiter.join([""] + src.ids))
        fh.write("\n")
        for r0, r1, rows in src.blocks(block_rows):
            text = np.asarray(rows, dtype=str)
            fh.write("".join("%s%s%s\n" % (src.ids[r0 + k], delimiter, delimiter.join(text[k]))
                             for k in range(r1 - r0)))
    finally:
        if own:
            fh.close()
    return src.n


def write_binary_dm(source, path, ids=None, block_rows=1024):
    """Write the distances as the reference's ``binary_dm`` HDF5 layout, row block by row block."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is absent from the build image
        raise ImportError("write_binary_dm needs h5py (HDF5); it is not installed here") from e
    src = _RowSource(source, ids)
    with h5py.File(path, "w") as f:
        f.create_dataset("format", data=np.array([b"BDSM"]))      # _set_header, binary_dm.py:228-231
        f.create_dataset("version", data=np.array([b"2020.12"]))
        f.create_dataset("order", data=np.array([x.encode("utf-8") for x in src.ids]))
        mat = f.create_dataset("matrix", shape=(src.n, src.n), dtype=np.float64)
        for r0, r1, rows in src.blocks(block_rows):
            mat[r0:r1, :] = rows
    return src.n
