"""Consumers of the distance matrix that run on the engine ("next" rows of SURVEY.md section 8f).

``permanova`` follows /root/reference/skbio/stats/distance/_permanova.py:22-239: same input
handling and messages (``_preprocess_input_sng``, ``_df_to_vector``,
/root/reference/skbio/stats/distance/_base.py:2207-2337), same permutation stream
(``get_rng(seed)`` then one ``rng.permutation(grouping)`` per permutation,
_base.py:2340-2364), same result Series (``_build_results``).  Only the arithmetic moves: the
pseudo-F partial statistic of the observed grouping and of all permutations is computed in
one GPU pass over the condensed distances (``b2d_permanova_sw``).
"""

from __future__ import annotations

import numpy as np

from . import _capi


def _get_rng(seed):
    # /root/reference/skbio/util/_random.py:89-118
    if isinstance(seed, np.random.Generator):
        return seed
    if seed is None:
        seed = np.random.randint(np.iinfo(np.int32).max + 1)
    elif isinstance(seed, np.random.RandomState):
        seed = seed.randint(np.iinfo(np.int32).max + 1)
    elif not np.issubdtype(type(seed), np.integer):
        raise ValueError("Invalid seed. It must be an integer or a random generator instance.")
    return np.random.default_rng(seed)


def _preprocess_input_sng(ids, sample_size, grouping, column):
    import pandas as pd

    def df_to_vector(df, col):
        if col not in df:
            raise ValueError("Column '%s' not in DataFrame." % col)
        g = df.reindex(ids, axis=0).loc[:, col]
        if g.isnull().any():
            raise ValueError("One or more IDs in the distance matrix are not in the data frame.")
        return g.tolist()

    if isinstance(grouping, pd.DataFrame):
        if column is None:
            raise ValueError("Must provide a column name if supplying a DataFrame.")
        grouping = df_to_vector(grouping, column)
    elif isinstance(grouping, pd.Series):
        if (column is not None) and (column != grouping.name):
            raise ValueError(
                "Column name does not match your Series name. Try not providing column at all.")
        grouping = df_to_vector(grouping.to_frame(), grouping.name)
    elif column is not None:
        raise ValueError("Must provide a DataFrame if supplying a column name.")
    if len(grouping) != sample_size:
        raise ValueError(
            "Grouping vector size must match the number of IDs in the distance matrix.")
    groups, grouping = np.unique(grouping, return_inverse=True)
    num_groups = len(groups)
    if num_groups == len(grouping):
        raise ValueError(
            "All values in the grouping vector are unique. This method cannot "
            "operate on a grouping vector with only unique values (e.g., "
            "there are no 'within' distances because each group of objects "
            "contains only a single object).")
    if num_groups == 1:
        raise ValueError(
            "All values in the grouping vector are the same. This method "
            "cannot operate on a grouping vector with only a single group of "
            "objects (e.g., there are no 'between' distances because there is "
            "only a single group).")
    return num_groups, grouping


def permanova(distmat, grouping, column=None, permutations=999, seed=None, device=0):
    """PERMANOVA pseudo-F and permutation p-value (drop-in for skbio.stats.distance.permanova)."""
    import pandas as pd

    if not (hasattr(distmat, "condensed_form") and hasattr(distmat, "ids")):
        raise TypeError("Input must be a DistanceMatrix.")
    sample_size = distmat.shape[0]
    num_groups, grouping = _preprocess_input_sng(distmat.ids, sample_size, grouping, column)
    if permutations < 0:
        raise ValueError("Number of permutations must be greater than or equal to zero.")
    if _capi.load().b2d_device_count() <= 0:
        raise _capi.EngineError("no CUDA device visible: the B200 engine has no CPU fallback.")
    group_sizes = np.bincount(grouping)
    rng = _get_rng(seed)
    rows = [grouping]
    for _ in range(permutations):
        rows.append(rng.permutation(grouping))
    # one extra row puts everything in one group of weight 1/n: that row's s_W is s_T
    G = np.vstack(rows + [np.full(sample_size, num_groups)]).astype(np.int32)
    inv = np.concatenate([1.0 / group_sizes, [1.0 / sample_size]])
    s = _capi.permanova_sw(distmat.condensed_form(), G, inv, device=device)
    s_T, s_W = s[-1], s[:-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        f = ((s_T - s_W) / (num_groups - 1)) / (s_W / (sample_size - num_groups))
    stat = f[0]
    p_value = np.nan
    if permutations > 0:
        p_value = ((f[1:] >= stat).sum() + 1) / (permutations + 1)
    return pd.Series(
        data=["PERMANOVA", "pseudo-F", sample_size, num_groups, stat, p_value, permutations],
        index=["method name", "test statistic name", "sample size", "number of groups",
               "test statistic", "p-value", "number of permutations"],
        name="PERMANOVA results")
