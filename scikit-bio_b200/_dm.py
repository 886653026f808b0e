"""Result type of ``beta_diversity``.

When scikit-bio is importable the real ``skbio.stats.distance.DistanceMatrix`` is returned.
Otherwise (GPU boxes without scikit-bio) a minimal stand-in with the same constructor
contract for this path (/root/reference/skbio/stats/distance/_base.py:1961-2060, 1268-1302):
1-D input is a SciPy condensed vector, ids default to '0'..'n-1', duplicate ids or a length
mismatch raise, 1-D data are not checked for symmetry/hollowness (so NaN from Bray-Curtis
0/0 survives), ``condensed=True`` keeps the vector instead of expanding to n x n.
"""

from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from skbio.stats.distance import DistanceMatrix as _SkbioDistanceMatrix  # type: ignore
    from skbio.stats.distance import DistanceMatrixError as _SkbioDMError  # type: ignore
except Exception:  # pragma: no cover
    _SkbioDistanceMatrix = None
    _SkbioDMError = None


class DistanceMatrixError(Exception):
    """Mirror of skbio.stats.distance.DistanceMatrixError for the stand-in class."""


def _vec_to_shape(vec):
    # n(n-1)/2 = len  ->  n
    ln = len(vec)
    n = int(round((1.0 + np.sqrt(1.0 + 8.0 * ln)) / 2.0))
    if n * (n - 1) // 2 != ln:
        raise DistanceMatrixError(
            "Incompatible vector size. It must be a binomial coefficient n choose 2 for some "
            "integer n >= 2.")
    return n


def squareform(vec, n):
    out = np.zeros((n, n), dtype=np.float64)
    iu = np.triu_indices(n, k=1)
    out[iu] = vec
    out[(iu[1], iu[0])] = vec
    return out


class SimpleDistanceMatrix:
    """Small stand-in for skbio's DistanceMatrix (construct, index, condensed/redundant)."""

    def __init__(self, data, ids=None, validate=True, condensed=False):
        data = np.asarray(data, dtype=np.float64)
        if data.ndim == 1:
            n = 1 if data.size == 0 else _vec_to_shape(data)
        elif data.ndim == 2:
            if data.shape[0] != data.shape[1]:
                raise DistanceMatrixError("Data must be square (i.e., have the same number of "
                                          "rows and columns).")
            n = data.shape[0]
        else:
            raise DistanceMatrixError("Data must have exactly 1 or 2 dimensions.")
        if ids is None:
            ids = tuple(str(i) for i in range(n))
        else:
            ids = tuple(ids)
        if validate:
            if len(set(ids)) != len(ids):
                seen, dup = set(), []
                for i in ids:
                    if i in seen and i not in dup:
                        dup.append(i)
                    seen.add(i)
                raise DistanceMatrixError(
                    "IDs must be unique. Found the following duplicate IDs: %s"
                    % ", ".join(repr(e) for e in dup))
            if len(ids) == 0:
                raise DistanceMatrixError("IDs must be at least 1 in size.")
            if len(ids) != n:
                raise DistanceMatrixError(
                    "The number of IDs (%d) must match the number of rows/columns in the "
                    "data (%d)." % (len(ids), n))
        self._ids = ids
        self._index = {i: k for k, i in enumerate(ids)}
        self._n = n
        if data.ndim == 1:
            self._condensed = data
            self._square = None if condensed else squareform(data, n)
        else:
            self._square = data
            self._condensed = None

    @property
    def ids(self):
        return self._ids

    @property
    def shape(self):
        return (self._n, self._n)

    @property
    def data(self):
        if self._square is None:
            self._square = squareform(self._condensed, self._n)
        return self._square

    def condensed_form(self):
        if self._condensed is None:
            iu = np.triu_indices(self._n, k=1)
            self._condensed = np.ascontiguousarray(self._square[iu])
        return self._condensed

    def redundant_form(self):
        return self.data

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and all(k in self._index for k in key):
            return self.data[self._index[key[0]], self._index[key[1]]]
        if key in self._index:
            return self.data[self._index[key]]
        return self.data[key]

    def __eq__(self, other):
        return (isinstance(other, SimpleDistanceMatrix) and self._ids == other._ids
                and np.array_equal(self.data, other.data, equal_nan=False))

    def __ne__(self, other):
        return not self == other

    def __repr__(self):
        return f"<SimpleDistanceMatrix {self._n}x{self._n}>"


def make_distance_matrix(vec, ids, condensed=False, n=None):
    """``DistanceMatrix(distances, ids)`` as _driver.py:355 does, on whichever class exists."""
    if _SkbioDistanceMatrix is not None:
        if condensed:
            return _SkbioDistanceMatrix(vec, ids, condensed=True)
        if n == 1 and len(vec) == 0:
            return _SkbioDistanceMatrix(np.zeros((1, 1)), ids)
        return _SkbioDistanceMatrix(vec, ids)
    return SimpleDistanceMatrix(vec, ids, condensed=condensed)


DistanceMatrix = _SkbioDistanceMatrix if _SkbioDistanceMatrix is not None else SimpleDistanceMatrix
