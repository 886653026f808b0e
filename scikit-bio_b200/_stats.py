"""Consumers of the distance matrix that run on the engine ("next" rows of SURVEY.md section 8f).

* ``permanova`` - /root/reference/skbio/stats/distance/_permanova.py:22-239.  The pseudo-F partial
  statistic of the observed grouping and of every permutation is computed in ONE GPU pass over
  the condensed distances (``b2d_permanova_sw``) instead of one pass per permutation
  (/root/reference/skbio/stats/distance/_base.py:2340-2364).  The permutation stream is the
  reference's (one ``Generator.permutation(grouping)`` per permutation), so p-values are
  identical for the same seed.
* ``center_distance_matrix`` - the first step of ``pcoa``
  (/root/reference/skbio/stats/ordination/_cutils.pyx:19-123, ``e_matrix_means_cy`` +
  ``f_matrix_inplace_cy``): Gower centering of the condensed distances on the device, the n x n
  result streamed back in row blocks.

Input validation is scikit-bio's own when scikit-bio is importable (the engine then raises
exactly what the reference raises); otherwise a small stand-alone check with the same
conditions.
"""

from __future__ import annotations

import numpy as np

from . import _capi

_ALL_UNIQUE = ("All values in the grouping vector are unique. This method cannot operate on a "
               "grouping vector with only unique values (e.g., there are no 'within' distances "
               "because each group of objects contains only a single object).")
_ALL_SAME = ("All values in the grouping vector are the same. This method cannot operate on a "
             "grouping vector with only a single group of objects (e.g., there are no 'between' "
             "distances because there is only a single group).")


def _rng_from(seed):
    """The Generator ``skbio.util.get_rng`` would return for `seed`."""
    try:  # pragma: no cover - depends on the environment
        from skbio.util import get_rng  # type: ignore

        return get_rng(seed)
    except ImportError:
        pass
    if isinstance(seed, np.random.Generator):
        return seed
    if isinstance(seed, np.random.RandomState):
        return np.random.default_rng(seed.randint(2 ** 31))
    if seed is None:
        return np.random.default_rng(np.random.randint(2 ** 31))
    if isinstance(seed, (int, np.integer)):
        return np.random.default_rng(seed)
    raise ValueError("Invalid seed. It must be an integer or a random generator instance.")


def _group_codes(ids, grouping, column):
    """``(number of groups, integer code per object)`` for a grouping given as a sequence, a
    pandas Series or a DataFrame column, aligned to `ids`."""
    try:  # pragma: no cover - depends on the environment
        from skbio.stats.distance._base import _preprocess_input_sng  # type: ignore

        return _preprocess_input_sng(ids, len(ids), grouping, column)
    except ImportError:
        pass
    import pandas as pd

    labelled = None   # a Series indexed by object id, when the caller gave one
    if isinstance(grouping, pd.DataFrame):
        if column is None:
            raise ValueError("Must provide a column name if supplying a DataFrame.")
        if column not in grouping:
            raise ValueError("Column '%s' not in DataFrame." % column)
        labelled = grouping[column]
    elif isinstance(grouping, pd.Series):
        if column is not None and column != grouping.name:
            raise ValueError(
                "Column name does not match your Series name. Try not providing column at all.")
        labelled = grouping
    elif column is not None:
        raise ValueError("Must provide a DataFrame if supplying a column name.")
    if labelled is not None:
        aligned = labelled.reindex(list(ids))
        if aligned.isnull().any():
            raise ValueError("One or more IDs in the distance matrix are not in the data frame.")
        grouping = aligned.tolist()
    if len(grouping) != len(ids):
        raise ValueError("Grouping vector size must match the number of IDs in the distance matrix.")
    labels, codes = np.unique(grouping, return_inverse=True)
    if len(labels) == len(codes):
        raise ValueError(_ALL_UNIQUE)
    if len(labels) == 1:
        raise ValueError(_ALL_SAME)
    return len(labels), codes


def permanova(distmat, grouping, column=None, permutations=999, seed=None, device=0):
    """PERMANOVA pseudo-F and permutation p-value (drop-in for skbio.stats.distance.permanova)."""
    import pandas as pd

    if not (hasattr(distmat, "condensed_form") and hasattr(distmat, "ids")):
        raise TypeError("Input must be a DistanceMatrix.")
    n = distmat.shape[0]
    k, codes = _group_codes(distmat.ids, grouping, column)
    if permutations < 0:
        raise ValueError("Number of permutations must be greater than or equal to zero.")
    if _capi.load().b2d_device_count() <= 0:
        raise _capi.EngineError("no CUDA device visible: the B200 engine has no CPU fallback.")
    rng = _rng_from(seed)
    # row 0: the observed grouping; rows 1..P: its permutations; last row: every object in one
    # extra group of weight 1/n, whose s_W is the total sum of squares s_T
    table = np.empty((permutations + 2, n), dtype=np.int32)
    table[0] = codes
    for r in range(1, permutations + 1):
        table[r] = rng.permutation(codes)
    table[-1] = k
    weights = np.append(1.0 / np.bincount(codes, minlength=k), 1.0 / n)
    s = _capi.permanova_sw(distmat.condensed_form(), table, weights, device=device)
    s_total, s_within = s[-1], s[:-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        pseudo_f = ((s_total - s_within) / (k - 1)) / (s_within / (n - k))
    p_value = np.nan
    if permutations > 0:
        p_value = (np.count_nonzero(pseudo_f[1:] >= pseudo_f[0]) + 1) / (permutations + 1)
    return pd.Series(
        ["PERMANOVA", "pseudo-F", n, k, pseudo_f[0], p_value, permutations],
        index=["method name", "test statistic name", "sample size", "number of groups",
               "test statistic", "p-value", "number of permutations"],
        name="PERMANOVA results")


def center_distance_matrix(distmat, device=0, return_means=False):
    """Gower-centred matrix ``F`` of a distance matrix (Legendre & Legendre eq. 9.20-9.21): the
    first step of ``skbio.stats.ordination.pcoa``, on the GPU.

    `distmat` is a DistanceMatrix (anything with ``condensed_form()``) or a 1-D condensed vector.
    Returns the n x n float64 array ``F[i, j] = E[i, j] - mean_i - mean_j + mean`` with
    ``E = -d^2 / 2``; with ``return_means`` also the row means of ``E`` and its global mean.
    """
    vec = distmat.condensed_form() if hasattr(distmat, "condensed_form") else np.asarray(distmat)
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    if vec.ndim != 1:
        raise ValueError("center_distance_matrix takes a DistanceMatrix or a condensed vector.")
    f, row_means, global_mean = _capi.center_condensed(vec, device=device)
    return (f, row_means, global_mean) if return_means else f
