"""CPU oracle for the multi-class branch of PHet ``fit_predict`` (K >= 3 conditions, delta_type="iqr",
group_test="anova", no pseudobulk) -- SURVEY.md 8(f) item 3.

TEST INFRASTRUCTURE ONLY (see oracle/phet_oracle.py's header: same rules).  Restates
``/root/reference/src/model/phet.py:301-306`` (combinations), ``309-332`` (subsample size), ``348-382``
(per-subsample one-way ANOVA), ``385-417`` (delta-IQR per combination), ``435-444`` (aggregation),
``449-462`` (Fisher), ``465-498`` (Step 3 per combination over min_size) and ``500-516`` stage by stage, exposing
the intermediates.  Labels must already be 0..K-1 (the reference mixes pre- and post-LabelEncoder labels,
phet.py:214,303,324 vs 220).

Pinning: ``tests/golden/make_golden_multiclass.py`` runs the UNMODIFIED reference under oracle/ref_shim.py on the
same seed and REQUIRES max|oracle - reference| == 0.0 before writing the fixtures (tests/golden/multi_*.npz).
"""
from __future__ import annotations

from itertools import combinations

import numpy as np
from scipy import stats as _stats

from . import phet_oracle as po


def combos_of(K: int, strategy: str):
    """phet.py:301-306."""
    if strategy == "one_vs_all":
        return list(zip(range(K), [0] * K))
    return list(combinations(range(K), 2))


def combo_members(y, i, j, strategy):
    """phet.py:389-392 / 476-480: rows of the two sides of a combination."""
    ei = np.where(y == i)[0]
    ej = np.where(y != i)[0] if strategy == "one_vs_all" else np.where(y == j)[0]
    return ei, ej


def subsample_size_multi(y, K, subsampling_size, strategy):
    """phet.py:309-332."""
    sizes = [int(np.sum(y == c)) for c in range(K)]
    if subsampling_size == "optimum":
        out = []
        for i, j in combos_of(K, strategy):
            ei, ej = combo_members(y, i, j, strategy)
            out.append(po.optimum_sample_size(len(ei), len(ej), alpha=0.05, margin_error=0.1))
        return int(np.min(out))
    if subsampling_size == "sqrt":
        mn = np.min(sizes)
        if mn <= 2 or int(np.sqrt(mn)) < 8:
            return int(mn)
        return int(np.sqrt(mn))
    return int(subsampling_size)


def anova_p_plain(groups):
    """scipy.stats.f_oneway(*groups) (axis 0, equal_var=True), formula restatement in the groups' dtype:
    sums of squares about the pooled mean, F -> special.fdtrc; constant-input rules as in _stats_py.py."""
    from scipy import special
    alldata = np.concatenate(groups, axis=0)
    bign = alldata.shape[0]
    K = len(groups)
    offset = alldata.mean(axis=0, keepdims=True)
    centred = alldata - offset
    nss = centred.sum(axis=0) ** 2.0 / bign
    sstot = np.einsum("ij,ij->j", centred, centred) - nss
    ssbn = 0
    for gsamp in groups:
        ssbn = ssbn + (gsamp - offset).sum(axis=0) ** 2.0 / gsamp.shape[0]
    ssbn = ssbn - nss
    sswn = sstot - ssbn
    with np.errstate(divide="ignore", invalid="ignore"):
        f = (ssbn / (K - 1)) / (sswn / (bign - K))
    is_const = np.stack([np.all(np.diff(gsamp, axis=0) == 0, axis=0) for gsamp in groups])
    f = np.where(np.all(is_const, axis=0), np.inf, f)
    f = np.where(np.all(np.diff(alldata, axis=0) == 0, axis=0), np.nan, f)
    return special.fdtrc(K - 1, bign - K, f)


def kruskal_p_plain(groups):
    """scipy.stats.kruskal(*groups) along axis 0, formula restatement: average ranks of the pooled values,
    H = 12/(n(n+1)) sum_k R_k^2/n_k - 3(n+1), tie correction 1 - sum(t^3 - t)/(n^3 - n), p = chi2.sf(H, K-1)."""
    alldata = np.concatenate(groups, axis=0)
    n, G = alldata.shape
    K = len(groups)
    H = np.empty(G)
    for g in range(G):
        col = alldata[:, g]
        ranked = _stats.rankdata(col)
        _, cnt = np.unique(col, return_counts=True)
        ties = 1 - np.sum(cnt.astype(np.float64) ** 3 - cnt) / (float(n) ** 3 - n)
        ssbn, j = 0.0, 0
        for gs in groups:
            ssbn += ranked[j:j + gs.shape[0]].sum() ** 2 / gs.shape[0]
            j += gs.shape[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            H[g] = (12.0 / (n * (n + 1)) * ssbn - 3 * (n + 1)) / ties
    return _stats.chi2.sf(H, K - 1), H


def fit_predict_multiclass_oracle(X, y, normalize="zscore", percentiles=(25, 75), num_subsamples=1000,
                                  subsampling_size="sqrt", feature_weight=None, strategy="one_vs_all",
                                  aggregation="max", rng=np.random, library=True, group_test="anova"):
    """Replay of PHet.fit_predict for K >= 3 classes, consuming ``rng`` in the reference's order:
    S x K ANOVA draws; per combination S x (i draw, j draw); per gene, per combination (i perm, j perm).
    Returns a dict of every stage."""
    X = np.asarray(X)
    y = np.asarray(y)
    K = len(np.unique(y))
    assert K >= 3 and np.array_equal(np.unique(y), np.arange(K)), "labels must be 0..K-1, K >= 3"
    if feature_weight is None:
        feature_weight = [0.4, 0.3, 0.2, 0.1]
    fw = np.array(feature_weight) / np.sum(feature_weight)
    if normalize == "zscore":
        Xz = po.zscore_nan_to_num(X)
    else:
        Xz = np.array(X, copy=True)
        np.nan_to_num(Xz, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    G = X.shape[1]
    S = num_subsamples
    combos = combos_of(K, strategy)
    m = subsample_size_multi(y, K, subsampling_size, strategy)
    rows = [np.where(y == c)[0] for c in range(K)]
    sizes = [len(r) for r in rows]
    # ---- Step 1a: ANOVA over the K classes (phet.py:348-382)
    replace = any(m > n for n in sizes)
    idx_anova = np.empty((S, K, m), dtype=np.int64)
    P = np.zeros((G, S))
    for s in range(S):
        groups = []
        for c in range(K):
            t = rng.choice(a=range(sizes[c]), size=m, replace=replace)
            idx_anova[s, c] = rows[c][t]
            groups.append(Xz[idx_anova[s, c]])
        if group_test == "kruskal" and m > 5:          # phet.py:372-377
            p = _stats.kruskal(*groups)[1] if library else kruskal_p_plain(groups)[0]
        elif library:
            _, p = _stats.f_oneway(*groups)
        else:
            p = anova_p_plain(groups)
        p = np.array(p, copy=True)
        np.nan_to_num(p, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
        P[:, s] = p
    # ---- Step 1b: delta-IQR per combination (phet.py:385-417)
    C = len(combos)
    R = np.zeros((G, S, C))
    idx_combo = []
    for ci, (i, j) in enumerate(combos):
        ei, ej = combo_members(y, i, j, strategy)
        rep = bool(m > len(ei) or m > len(ej))
        ii = np.empty((S, m), dtype=np.int64)
        jj = np.empty((S, m), dtype=np.int64)
        for s in range(S):
            ii[s] = rng.choice(a=ei, size=m, replace=rep)
            jj[s] = rng.choice(a=ej, size=m, replace=rep)
            if library:
                a = _stats.iqr(Xz[ii[s]], axis=0, rng=percentiles, scale=1.0)
                b = _stats.iqr(Xz[jj[s]], axis=0, rng=percentiles, scale=1.0)
            else:
                a = po.iqr_plain(Xz[ii[s]], percentiles)
                b = po.iqr_plain(Xz[jj[s]], percentiles)
            d = np.absolute(a - b)
            np.nan_to_num(d, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
            R[:, s, ci] = d
        idx_combo.append((ii, jj))
    rsum = R.sum(axis=1)                                   # (G, C)
    r = np.mean(R, axis=1)
    r = np.mean(r, axis=1) if aggregation == "mean" else np.max(r, axis=1)
    np.nan_to_num(r, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    # ---- Step 2
    fsum = np.sum(po.fisher_terms(P), axis=1)
    f = po.fisher_standardise(fsum, S)
    # ---- Step 3 (phet.py:465-498): slices over [0, min_size) for every combination
    min_size = int(np.min(sizes))
    members = [combo_members(y, i, j, strategy) for i, j in combos]
    # phet.py:488-492: slice starts run over [0, min_size) of ALL classes, but the last slice takes
    # min(len(column i), len(column j)) - start elements -- with K >= 3 that can exceed m
    slices_c = []
    for ei, ej in members:
        sl = []
        for start in range(0, min_size, m):
            n = m
            if start + m >= min_size:
                n = min(len(ei), len(ej)) - start
            sl.append((start, n))
        slices_c.append(sl)
    slices = slices_c[0]
    tabs = {}
    for sl in slices_c:
        for _, n in sl:
            if n not in tabs:
                tabs[n] = po.ks_table(n)
    perms = [[None, None] for _ in combos]
    for ci, (i, j) in enumerate(combos):
        ei, ej = combo_members(y, i, j, strategy)
        perms[ci][0] = np.empty((G, len(ei)), dtype=np.int32)
        perms[ci][1] = np.empty((G, len(ej)), dtype=np.int32)
    H = np.zeros((G, C, len(slices)), dtype=np.int32)
    o_combo = np.zeros((G, C))
    for g in range(G):
        for ci in range(C):
            ei, ej = members[ci]
            pi = rng.permutation(len(ei))
            pj = rng.permutation(len(ej))
            perms[ci][0][g] = pi
            perms[ci][1][g] = pj
            col_i = Xz[ei, g][pi]
            col_j = Xz[ej, g][pj]
            ps = []
            for k, (st, n) in enumerate(slices_c[ci]):
                h = po.ks_h_equal_n(col_i[st:st + n], col_j[st:st + n])
                H[g, ci, k] = h
                ps.append(tabs[n][h])
            o_combo[g, ci] = np.min(ps)
    o = np.zeros((G, 1))
    for g in range(G):
        o[g] = np.mean(list(o_combo[g]))
    o = o[:, 0]
    bins, edges = po.kbins_uniform(o, len(fw))
    scores = po.combine(r, f, bins, fw)
    return dict(scores=scores, m=m, K=K, combos=combos, idx_anova=idx_anova, idx_combo=idx_combo, P=P, fsum=fsum,
                rsum=rsum, r=r, f=f, min_size=min_size, slices=slices, slices_c=slices_c, perms=perms, H=H, o_combo=o_combo, o=o,
                bins=bins, edges=edges, bin_counts=np.bincount(bins, minlength=len(fw)).astype(np.int64), Xz=Xz)
