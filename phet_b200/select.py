"""What consumes the scores: selection by a fitted gamma tail and descending ranking.

Arithmetic unchanged from ``/root/reference/src/utility/utils.py:80-136`` (``significant_features`` /
``sort_features``, the X.shape[1] == 1 branch used for PHet).  These produce the integer outputs the
north star requires to be bit-exact (selected-feature indices and top-k ranking).  The gamma MLE and
its tail test stay scipy calls on the (G, 1) score vector; the ordering can
come from the device (``phet_rank`` -> ``PHet.ranking_``): pass it as ``order=`` and no pandas sort runs.
"""
from __future__ import annotations

from typing import Literal

import numpy as np
import pandas as pd
from scipy.stats import expon, gamma, lognorm, scoreatpercentile, zscore


def selected_indices(X: np.ndarray, alpha: float = 0.05,
                     fit_type: Literal["expon", "gamma", "lognormal", "percentile"] = "gamma", per: int = 95):
    """Indices kept by ``significant_features`` (utils.py:83-99)."""
    tempX = np.copy(X)
    if X.shape[1] != 1:
        tempX = X[:, 3]
    tempX += 0.001
    if fit_type == "percentile":
        temp = scoreatpercentile(tempX, per, interpolation_method="higher")
        return np.where(tempX > temp)[0]
    if fit_type == "lognormal":
        shape, loc, scale = lognorm.fit(tempX)
        return np.where((1 - lognorm.cdf(tempX, shape, loc=loc, scale=scale)) < alpha)[0]
    if fit_type == "expon":
        shape, scale = expon.fit(tempX)
        return np.where((1 - expon.cdf(tempX, shape, scale=scale)) < alpha)[0]
    shape, loc, scale = gamma.fit(zscore(tempX))
    return np.where((1 - gamma.cdf(zscore(tempX), shape, loc=loc, scale=scale)) < alpha)[0]


def sort_features(X: np.ndarray, features_name, ascending: bool = False, order=None, **_ignored) -> pd.DataFrame:
    """utils.py:107-136 for a (G, 1) score matrix.  ``file_name=`` / ``save_path=`` (passed by
    src/train.py:496-504 although the reference function does not take them) are accepted and ignored.
    ``order``: descending order of the rows of X computed on the device (``phet_rank``); rows are then
    gathered in that order instead of being sorted by pandas."""
    if X.shape[1] != 1:
        raise Exception("Please provide correct shape for X")
    if order is not None:
        order = np.asarray(order, dtype=np.int64)
        if ascending:
            order = order[::-1]
        return pd.DataFrame({"features": np.asarray(features_name, dtype=object)[order],
                             "score": np.asarray(X)[order, 0]})
    df = pd.concat([pd.DataFrame(features_name), pd.DataFrame(X)], axis=1)
    df.columns = ["features", "score"]
    return df.sort_values(by=["score"], ascending=ascending).reset_index(drop=True)


def significant_features(X: np.ndarray, features_name, alpha: float = 0.05, fit_type="gamma", per: int = 95,
                         order=None, **_ignored) -> pd.DataFrame:
    """utils.py:80-104.  ``order``: device ranking of ALL rows of X (``PHet.ranking_``); the selected rows keep
    their relative order in it, which is their descending order."""
    sel = selected_indices(X, alpha=alpha, fit_type=fit_type, per=per)
    if order is not None:
        order = np.asarray(order, dtype=np.int64)
        if len(sel) != 0:
            keep = np.zeros(X.shape[0], dtype=bool)
            keep[sel] = True
            order = order[keep[order]]
        return sort_features(X=X, features_name=features_name, order=order)
    if len(sel) != 0:
        X = X[sel]
        features_name = np.array(features_name)[sel].tolist()
    return sort_features(X=X, features_name=features_name)


def top_k(X: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k highest scores, in the order pandas' descending sort_values yields."""
    df = pd.DataFrame({"idx": np.arange(X.shape[0]), "score": X[:, 0]})
    return df.sort_values(by=["score"], ascending=False)["idx"].to_numpy()[:k]
