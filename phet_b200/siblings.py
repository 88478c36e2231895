"""Drop-in replacements for the reference's two sibling IQR scorers.

``HIQR`` mirrors ``src/model/hvf.py:59-96`` and ``DeltaIQRMean`` mirrors
``src/model/deltaiqrmean.py:16-85``: same constructor arguments, same ``fit_predict(X, y,
control_class, case_class)`` signature, same output shapes and columns.  The per-gene Python loops of
the reference (``scipy.stats.iqr`` / ``np.mean`` / ``scipy.stats.ttest_ind`` on whole classes) are one
CUDA kernel here (``phet_class_stats``: exact radix select of the order statistics + two-pass moments
per gene, csrc/classstats.cu); the few vector operations after it stay on the host in the reference's
order.  Binary labels only; ``normalize`` in {"zscore", None} (as for PHet).
There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _capi, hostmath


def _device_class_stats(X, y, normalize, iqr_range, device=0):
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    X = np.ascontiguousarray(X)
    y = np.asarray(y)
    num_classes = len(np.unique(y))
    if num_classes != 2:
        raise NotImplementedError("phet_b200 covers two conditions (labels 0 and 1); got %d classes" % num_classes)
    if normalize not in ("zscore", None):
        raise NotImplementedError("normalize=%r is outside the accelerated path (use 'zscore' or None)" % (normalize,))
    if len(y) != X.shape[0]:
        raise Exception("The number of labels does not match the number of examples!")
    examples_i = np.where(y == 0)[0]   # the reference indexes classes by value: np.where(y == i), i in range(k)
    examples_j = np.where(y == 1)[0]
    n0, n1 = len(examples_i), len(examples_j)
    if n0 < 1 or n1 < 1:
        raise Exception("labels must be 0 and 1")
    qdt = np.float32 if X.dtype == np.float32 else np.float64   # scipy.stats.iqr works in X's dtype
    q3 = []
    for n in (n0, n1, n0 + n1):
        (j0, k0, g0), (j1, k1, g1) = hostmath.quantile_positions(n, iqr_range, dtype=qdt)
        q3.append(_capi.QuantilePos(j0, k0, j1, k1, g0, g1))
    ctx = _capi.Context(int(device))
    try:
        ctx.set_classes(examples_i, examples_j)
        ctx.load_matrix(X, normalize=_capi.NORM_ZSCORE if normalize == "zscore" else _capi.NORM_NONE)
        out = ctx.class_stats(q3)
        launches = ctx.launch_count()
    finally:
        ctx.close()
    # iqr columns hold values of X's dtype (the kernel forms them in that dtype, like scipy.stats.iqr)
    return out, n0, n1, launches, X.dtype


class HIQR:
    """``hvf.py:59-96``: IQR of every feature over all cells, or summed over the conditions."""

    def __init__(self, per_condition: bool = False, normalize: str = "zscore", iqr_range=(25, 75), device: int = 0):
        self.per_condition = per_condition
        self.normalize = normalize
        self.iqr_range = iqr_range
        self.device = device

    def fit_predict(self, X, y, control_class: int = 0, case_class: int = 1):
        st, _, _, self.launches_, dt = _device_class_stats(X, y, self.normalize, self.iqr_range, self.device)
        self.class_stats_ = st
        results = st[0] + st[1] if self.per_condition else st[2].copy()   # hvf.py:86-93 (float64 accumulator)
        np.nan_to_num(results, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
        return np.reshape(results, (results.shape[0], 1))


class DeltaIQRMean:
    """``deltaiqrmean.py:16-85``: columns iqr, median_diff, ttest, score, class_diff."""

    def __init__(self, calculate_deltamean: bool = True, normalize: str = None, iqr_range=(25, 75), device: int = 0):
        self.calculate_deltamean = calculate_deltamean
        self.normalize = normalize
        self.iqr_range = iqr_range
        self.device = device

    def fit_predict(self, X, y, control_class: int = 0, case_class: int = 1):
        st, n0, n1, self.launches_, dt = _device_class_stats(X, y, self.normalize, self.iqr_range, self.device)
        self.class_stats_ = st
        iqr0, iqr1, _, mean0, mean1, ss0, ss1 = st
        temp_iqr = (iqr0.astype(dt) - iqr1.astype(dt)).astype(np.float64)   # deltaiqrmean.py:66-70 (X's dtype)
        temp_iqr[np.isnan(temp_iqr)] = 0
        temp_mean = mean0 - mean1                                # 71-75
        temp_mean[np.isnan(temp_mean)] = 0
        # scipy.stats.ttest_ind defaults (equal_var=True): pooled variance, df = n0 + n1 - 2   (76)
        with np.errstate(divide="ignore", invalid="ignore"):
            df = n0 + n1 - 2.0
            svar = (ss0 + ss1) / df
            statistic = (mean0 - mean1) / np.sqrt(svar * (1.0 / n0 + 1.0 / n1))
        classes = np.where(temp_iqr <= 0, 0, 1)                 # 81-85
        diff_iqrs = np.abs(temp_iqr)
        diff_means = np.abs(temp_mean)
        ttests = np.abs(statistic)
        with np.errstate(divide="ignore", invalid="ignore"):
            score = diff_iqrs / np.sum(diff_iqrs)
            if self.calculate_deltamean:
                score = score + diff_means / np.sum(diff_means)
        results = np.stack([diff_iqrs, diff_means, ttests, score, classes.astype(np.float64)], axis=1)
        np.nan_to_num(results, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
        return results
