"""CPU restatement of the reference's sibling IQR scorers (TEST INFRASTRUCTURE ONLY).

Nothing under ``phet_b200/`` may import this module; it is the checker for ``phet_b200.siblings``
(tests/, bench legs).  Pinned by execution: ``tests/golden/make_golden_siblings.py`` runs the
UNMODIFIED ``/root/reference/src/model/hvf.py`` (``HIQR``) and ``deltaiqrmean.py``
(``DeltaIQRMean``) and requires ``max|oracle - reference| == 0.0`` on every committed case.

The restatement exposes the per-class intermediates (IQR, mean, sum of squared deviations) that the
reference folds into its output columns, so the device kernel can be checked stage by stage.
"""
from __future__ import annotations

import numpy as np
from scipy import stats
from scipy.stats import iqr, zscore


def _normalize(X, normalize):
    # hvf.py:84-85 / deltaiqrmean.py:39-40: zscore(X, axis=0) with NO nan_to_num (constant genes -> NaN)
    if normalize == "zscore":
        return zscore(X, axis=0)
    if normalize is None:
        return X
    raise ValueError("oracle covers normalize in {'zscore', None}")


def class_stats(X, y, normalize, iqr_range):
    """Per gene: iqr(class 0), iqr(class 1), iqr(all), mean0, mean1, ss0, ss1 (X's dtype, like the reference)."""
    X = _normalize(np.asarray(X), normalize)
    i0, i1 = np.where(y == 0)[0], np.where(y == 1)[0]
    with np.errstate(invalid="ignore"):
        q0 = iqr(X[i0], axis=0, rng=iqr_range, scale=1.0)      # hvf.py:90, deltaiqrmean.py:63
        q1 = iqr(X[i1], axis=0, rng=iqr_range, scale=1.0)      # deltaiqrmean.py:64
        qa = iqr(X, axis=0, rng=iqr_range, scale=1.0)          # hvf.py:93
    A, B = np.ascontiguousarray(X[i0].T), np.ascontiguousarray(X[i1].T)   # one contiguous row per gene
    with np.errstate(invalid="ignore"):                                     # deltaiqrmean.py:71-72, per feature as there
        m0 = np.array([np.mean(X[i0, g]) for g in range(X.shape[1])])
        m1 = np.array([np.mean(X[i1, g]) for g in range(X.shape[1])])
    ss0 = ((A.astype(np.float64) - m0[:, None].astype(np.float64)) ** 2).sum(axis=1)
    ss1 = ((B.astype(np.float64) - m1[:, None].astype(np.float64)) ** 2).sum(axis=1)
    return dict(iqr0=q0, iqr1=q1, iqr_all=qa, mean0=m0, mean1=m1, ss0=ss0, ss1=ss1, Xn=X, i0=i0, i1=i1)


def hiqr(X, y, per_condition=False, normalize="zscore", iqr_range=(25, 75)):
    """hvf.py:66-96."""
    st = class_stats(X, y, normalize, iqr_range)
    results = np.zeros(X.shape[1])
    if per_condition:
        results += st["iqr0"]        # hvf.py:88-90: results[g] += iqr(class) for class 0 then class 1
        results += st["iqr1"]
    else:
        results[:] = st["iqr_all"]
    np.nan_to_num(results, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return results.reshape(-1, 1)


def delta_iqr_mean(X, y, calculate_deltamean=True, normalize=None, iqr_range=(25, 75)):
    """deltaiqrmean.py:22-103, two classes (one combination per feature)."""
    st = class_stats(X, y, normalize, iqr_range)
    Xn, i0, i1 = st["Xn"], st["i0"], st["i1"]
    G = X.shape[1]
    ttests = np.empty(G)
    with np.errstate(invalid="ignore", divide="ignore"):
        for g in range(G):                                      # deltaiqrmean.py:76 (per-feature call, as there)
            ttests[g] = stats.ttest_ind(Xn[i0, g], Xn[i1, g])[0]
    # the reference keeps per-feature python lists of numpy scalars (float32 stays float32, a NaN becomes the
    # python int 0), so the dtype of the score column follows numpy's promotion of those lists: restated as is
    diff_iqrs, diff_means, classes = [], [], []
    for g in range(G):
        ti = st["iqr0"][g] - st["iqr1"][g]
        if np.isnan(ti):
            ti = 0
        tm = st["mean0"][g] - st["mean1"][g]
        if np.isnan(tm):
            tm = 0
        iqrs, means = [ti], [tm]
        classes.append(0 if max(iqrs) <= 0 else 1)              # deltaiqrmean.py:81-85
        diff_iqrs.append(max(np.abs(iqrs)))                     # 88-89
        diff_means.append(max(np.abs(means)))
    classes = np.array(classes)
    with np.errstate(invalid="ignore", divide="ignore"):
        score = np.array(diff_iqrs) / np.sum(diff_iqrs)
        if calculate_deltamean:
            score = score + np.array(diff_means) / np.sum(diff_means)
    results = np.stack([np.array(diff_iqrs, dtype=np.float64), np.array(diff_means, dtype=np.float64), np.abs(ttests),
                        np.asarray(score, dtype=np.float64), classes.astype(np.float64)], axis=1)
    np.nan_to_num(results, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return results
