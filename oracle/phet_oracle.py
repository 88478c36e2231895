"""CPU oracle for the PHet ``fit_predict`` hot path (binary, delta_type="iqr", no pseudobulk).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as the checker
or the timed CPU baseline.  The product (``phet_b200/``) never imports it; the product path
fails loudly when the CUDA library is missing.

It restates ``/root/reference/src/model/phet.py:163-518`` stage by stage so that every
intermediate the reference hides (index sets, P[g,s], delta[g,s], KS h, o, bins) is exposed
for per-stage parity.  The arithmetic lives in third-party libraries the reference calls
(none vendored under /root/reference; ``requirements.txt:1-14`` gives lower bounds only):

* scipy (>=1.11; this image 1.18.1): ``zscore`` phet.py:291, ``iqr`` 413-414, ``ttest_ind`` 426,
  ``ks_2samp`` 493.  The oracle calls the same public functions for the "library" flavour and
  also restates their published formulas in plain numpy (``*_plain`` functions) so the device
  code has an arithmetic specification to follow.
* statsmodels (>=0.14; ABSENT here): ``ztest`` phet.py:428, restated in ``ztest_pooled``.
* scikit-learn (>=1.3; here 1.9.0): ``KBinsDiscretizer(uniform, ordinal)`` phet.py:500-501,
  restated in ``kbins_uniform`` (np.linspace edges + searchsorted side="right").
* numpy legacy global RNG: ``np.random.choice`` 401-402, ``np.random.permutation`` 486-487.

Pinning status: the reference ships NO tests or golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, executed in the build
container under ``oracle/ref_shim.py``; the generating script is ``tests/golden/make_golden.py``
and the vectors are committed under ``tests/golden/``.  ``tests/test_oracle_vs_golden.py``
checks ``fit_predict_oracle`` against them bit-exactly (max abs diff 0.0).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy import stats as _stats

SEED_VALUE = 0.001  # phet.py:29


# --------------------------------------------------------------------------------------
# Stage 0: normalisation (phet.py:291,299)
# --------------------------------------------------------------------------------------
def zscore_nan_to_num(X: np.ndarray) -> np.ndarray:
    """``zscore(X, axis=0)`` (population sd, ddof=0) then NaN/+-inf -> 0.  phet.py:291,299."""
    Xz = _stats.zscore(X, axis=0)
    np.nan_to_num(Xz, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return Xz


def zscore_plain(X: np.ndarray) -> np.ndarray:
    """Formula restatement of scipy.stats.zscore(axis=0, ddof=0): (x - mean) / sqrt(mean((x-mean)^2))."""
    X = np.asarray(X)
    mn = X.mean(axis=0, keepdims=True)
    sd = X.std(axis=0, ddof=0, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        Xz = (X - mn) / sd
    np.nan_to_num(Xz, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return Xz


# --------------------------------------------------------------------------------------
# Subsample size (phet.py:309-332) and index draws (phet.py:388-402)
# --------------------------------------------------------------------------------------
def optimum_sample_size(control_size: int, case_size: int, alpha: float = 0.05,
                        margin_error: float = 0.1) -> int:
    """phet.py:126-161."""
    total = control_size + case_size
    p = control_size / total
    q = 1 - p
    z = _stats.norm.ppf(1 - alpha / 2)
    numerator = (z ** 2 * p * q) / (margin_error ** 2)
    denominator = 1 + (numerator / total)
    return int(np.ceil(numerator / denominator))


def subsample_size(y: np.ndarray, subsampling_size="sqrt") -> int:
    """phet.py:309-332 for two classes labelled {0, 1}."""
    y = np.asarray(y)
    n0 = int(np.sum(y == 0))
    n1 = int(np.sum(y == 1))
    if subsampling_size == "optimum":
        return int(optimum_sample_size(n0, n1, alpha=0.05, margin_error=0.1))
    if subsampling_size == "sqrt":
        min_c = min(n0, n1)
        if min_c <= 2 or int(np.sqrt(min_c)) < 8:
            return int(min_c)
        return int(np.sqrt(min_c))
    return int(subsampling_size)


def draw_index_sets(y: np.ndarray, m: int, num_subsamples: int, rng=np.random):
    """Replay of phet.py:388-402: per subsample, control draw then case draw.

    Returns int64 arrays (S, m) of ORIGINAL row numbers, plus the ``replace`` flag.
    """
    y = np.asarray(y)
    examples_i = np.where(y == 0)[0]
    examples_j = np.where(y == 1)[0]
    replace = bool(m > len(examples_i) or m > len(examples_j))
    idx_i = np.empty((num_subsamples, m), dtype=np.int64)
    idx_j = np.empty((num_subsamples, m), dtype=np.int64)
    for s in range(num_subsamples):
        idx_i[s] = rng.choice(a=examples_i, size=m, replace=replace)
        idx_j[s] = rng.choice(a=examples_j, size=m, replace=replace)
    return idx_i, idx_j, replace


def draw_step3_perms(n0: int, n1: int, num_features: int, rng=np.random):
    """Replay of phet.py:486-487: per gene, permutation of the control column then the case
    column.  ``np.random.permutation(values)`` == ``values[np.random.permutation(len(values))]``
    with identical state advance (checked in tests), so positions are returned instead."""
    perm_i = np.empty((num_features, n0), dtype=np.int32)
    perm_j = np.empty((num_features, n1), dtype=np.int32)
    for g in range(num_features):
        perm_i[g] = rng.permutation(n0)
        perm_j[g] = rng.permutation(n1)
    return perm_i, perm_j


# --------------------------------------------------------------------------------------
# Stage 1: per-subsample statistics (phet.py:413-432)
# --------------------------------------------------------------------------------------
def quantile_positions(m: int, percentiles, dtype=np.float64):
    """Index arithmetic of scipy ``_quantile_hf`` method='linear' (_quantile.py:428-464) as
    reached from ``iqr`` (_stats_py.py:3187-3197): p = rng/100 cast to X's dtype,
    jg = p*n + (1-p); j = floor(jg)-1; g = jg % 1; both indices clipped to [0, n-1].

    Returns ((j_lo, jp1_lo, g_lo), (j_hi, jp1_hi, g_hi)) with g in ``dtype``.
    """
    lo, hi = percentiles
    rng = (lo / 100, hi / 100) if lo < hi else (hi / 100, lo / 100)
    out = []
    n = np.asarray(m, dtype=dtype)
    for p in rng:
        p = np.asarray(p, dtype=dtype)
        jg = p * n + (1 - p)
        jp1 = jg // 1
        j = jp1 - 1
        g = jg % 1
        if j < 0:
            g = np.asarray(0, dtype=dtype)
        j = np.clip(j, 0.0, n - 1)
        jp1 = np.clip(jp1, 0.0, n - 1)
        out.append((int(j), int(jp1), g.astype(dtype)[()]))
    return tuple(out)


def iqr_plain(A: np.ndarray, percentiles) -> np.ndarray:
    """Column-wise IQR, formula restatement of scipy.stats.iqr(A, axis=0, rng, scale=1.0)."""
    A = np.asarray(A)
    m = A.shape[0]
    (j0, k0, g0), (j1, k1, g1) = quantile_positions(m, percentiles, dtype=A.dtype)
    Y = np.sort(A, axis=0)
    q_lo = (1 - g0) * Y[j0] + g0 * Y[k0]
    q_hi = (1 - g1) * Y[j1] + g1 * Y[k1]
    return q_hi - q_lo


def pooled_stat_plain(A: np.ndarray, B: np.ndarray):
    """Pooled two-sample statistic in scipy's order (_stats_py.py:6815-6827, 6383-6398):
    v = var(ddof=1); svar = ((n1-1)v1 + (n2-1)v2)/df; denom = sqrt(svar*(1/n1+1/n2))."""
    n1, n2 = A.shape[0], B.shape[0]
    m1, m2 = A.mean(axis=0), B.mean(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        v1 = A.var(axis=0, ddof=1) if n1 > 1 else np.zeros_like(m1)
        v2 = B.var(axis=0, ddof=1) if n2 > 1 else np.zeros_like(m2)
        df = n1 + n2 - 2.0
        svar = ((n1 - 1) * v1 + (n2 - 1) * v2) / df
        denom = np.sqrt(svar * (1.0 / n1 + 1.0 / n2))
        t = (m1 - m2) / denom
    return t, df


def ttest_p_plain(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """p of scipy.stats.ttest_ind(A, B) defaults (equal_var=True): 2*stdtr(df, -|t|)."""
    from scipy import special
    t, df = pooled_stat_plain(A, B)
    return 2 * special.stdtr(np.asarray(df, dtype=t.dtype), -np.abs(t))


def ztest_pooled(A: np.ndarray, B: np.ndarray):
    """statsmodels.stats.weightstats.ztest(A, B) defaults, restated (see ref_shim.py).
    p is always float64 because scipy's norm.sf promotes."""
    n1, n2 = A.shape[0], B.shape[0]
    var = n1 * A.var(0) + n2 * B.var(0)
    var /= n1 + n2 - 2 * 1.0
    var *= 1.0 / n1 + 1.0 / n2
    std_diff = np.sqrt(var)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (A.mean(0) - B.mean(0) - 0) / std_diff
    p = _stats.norm.sf(np.abs(z)) * 2
    return z, p


def step1_subsample(Xz: np.ndarray, idx_i: np.ndarray, idx_j: np.ndarray, percentiles, m: int,
                    library: bool = True):
    """One iteration of the loop at phet.py:397-432.  Returns (p, delta), each (G,)."""
    A = Xz[idx_i]
    B = Xz[idx_j]
    if library:
        iq_i = _stats.iqr(A, axis=0, rng=percentiles, scale=1.0)
        iq_j = _stats.iqr(B, axis=0, rng=percentiles, scale=1.0)
    else:
        iq_i = iqr_plain(A, percentiles)
        iq_j = iqr_plain(B, percentiles)
    delta = np.absolute(iq_i - iq_j)
    np.nan_to_num(delta, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    if m < 30:
        if library:
            _, p = _stats.ttest_ind(A, B)
        else:
            p = ttest_p_plain(A, B)
    else:
        _, p = ztest_pooled(A, B)
    p = np.array(p, copy=True)
    np.nan_to_num(p, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return p, delta


def step1_all(Xz, idx_i, idx_j, percentiles, m, library=True):
    """phet.py:346,386,397-432: P (G,S) float64 and delta r (G,S) float64."""
    S = idx_i.shape[0]
    G = Xz.shape[1]
    P = np.zeros((G, S))
    R = np.zeros((G, S))
    for s in range(S):
        p, d = step1_subsample(Xz, idx_i[s], idx_j[s], percentiles, m, library=library)
        P[:, s] = p
        R[:, s] = d
    return P, R


# --------------------------------------------------------------------------------------
# Stage 2: Fisher (phet.py:449-462) and delta aggregate (phet.py:435-444)
# --------------------------------------------------------------------------------------
def fisher_terms(P: np.ndarray) -> np.ndarray:
    """-2 ln P with +inf (p == 0) replaced by 0.  phet.py:449-450."""
    with np.errstate(divide="ignore"):
        f = -2 * np.log(P)
    f[f == np.inf] = 0
    return f


def fisher_standardise(fsum: np.ndarray, num_subsamples: int) -> np.ndarray:
    """phet.py:458-462."""
    f = (fsum - 2 * num_subsamples) / np.sqrt(4 * num_subsamples)
    f = np.array(f, copy=True)
    f[f < 0] = SEED_VALUE
    f = np.absolute(f)
    np.nan_to_num(f, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return f


def delta_aggregate(R: np.ndarray) -> np.ndarray:
    """phet.py:435-444 with one combination: mean over subsamples (max over 1 combo)."""
    r = np.mean(R[:, :, None], axis=1)
    r = np.max(r, axis=1)
    np.nan_to_num(r, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return r


# --------------------------------------------------------------------------------------
# Stage 3: discriminative power (phet.py:466-498)
# --------------------------------------------------------------------------------------
def ks_h_equal_n(a: np.ndarray, b: np.ndarray) -> int:
    """Integer KS statistic for len(a)==len(b)==n: h = n*D = max over pooled values v of
    |#{a<=v} - #{b<=v}|  (scipy _stats_py.py:8053-8125 computes D; 7834-7836 rounds to h)."""
    a = np.sort(a)
    b = np.sort(b)
    allv = np.concatenate([a, b])
    ca = np.searchsorted(a, allv, side="right")
    cb = np.searchsorted(b, allv, side="right")
    return int(np.max(np.abs(ca - cb)))


def _prob_outside_square(n: int, h: int) -> float:
    """Pr(D_{n,n} >= h/n), two-sided, equal sizes.  Published closed form
    2*sum_k (-1)^(k-1) C(2n, n-kh)/C(2n, n), evaluated Horner-like exactly as scipy does
    (_stats_py.py:7713-7748) so the doubles are identical."""
    P = 0.0
    k = int(np.floor(n / h))
    while k >= 0:
        p1 = 1.0
        for j in range(h):
            p1 = (n - k * h - j) * p1 / (n + k * h + j + 1)
        P = p1 * (1.0 - P)
        k -= 1
    return 2 * P


def ks_table(n: int) -> np.ndarray:
    """p-value of ks_2samp (two-sided, method='auto', n1 == n2 == n <= 10000) as a function
    of the integer statistic h in 0..n.  Follows _ks_2samp_prob / _attempt_exact_2kssamp
    (_stats_py.py:8146-8180, 7824-7866): exact closed form; if it lands outside [0, 1]
    fall back to the asymptotic kstwo.sf(d, round(n/2)); clip to [0, 1]."""
    tab = np.empty(n + 1, dtype=np.float64)
    tab[0] = 1.0
    for h in range(1, n + 1):
        prob = _prob_outside_square(n, h)
        if not (0 <= prob <= 1):
            d = h * 1.0 / n
            en = float(n) * float(n) / (float(n) + float(n))
            prob = _stats.distributions.kstwo.sf(d, np.round(en))
        tab[h] = np.clip(prob, 0, 1)
    return tab


def step3_slices(min_c: int, m: int):
    """Slice (start, length) pairs of phet.py:488-492."""
    out = []
    for start in range(0, min_c, m):
        n = m
        if start + m >= min_c:
            n = min_c - start
        out.append((start, n))
    return out


def step3_h(Xz, ctrl, case, perm_i, perm_j, m):
    """Integer KS statistics per gene and slice: (G, n_slices) int32.  phet.py:473-495."""
    G = Xz.shape[1]
    min_c = min(len(ctrl), len(case))
    slices = step3_slices(min_c, m)
    H = np.zeros((G, len(slices)), dtype=np.int32)
    for g in range(G):
        col_i = Xz[ctrl, g][perm_i[g]]
        col_j = Xz[case, g][perm_j[g]]
        for k, (st, n) in enumerate(slices):
            H[g, k] = ks_h_equal_n(col_i[st:st + n], col_j[st:st + n])
    return H, slices


def step3_o_from_h(H: np.ndarray, slices) -> np.ndarray:
    """o[g] = min over slices of the exact KS p-value (phet.py:496,498), via tables."""
    tabs = {}
    for _, n in slices:
        if n not in tabs:
            tabs[n] = ks_table(n)
    cols = [tabs[n][H[:, k]] for k, (_, n) in enumerate(slices)]
    return np.min(np.stack(cols, axis=1), axis=1)


def step3_o_library(Xz, ctrl, case, perm_i, perm_j, m):
    """Same as the reference loop, calling scipy.stats.ks_2samp per slice (slow; small cases)."""
    G = Xz.shape[1]
    min_c = min(len(ctrl), len(case))
    o = np.zeros(G)
    for g in range(G):
        col_i = Xz[ctrl, g][perm_i[g]]
        col_j = Xz[case, g][perm_j[g]]
        ps = []
        for st, n in step3_slices(min_c, m):
            ps.append(_stats.ks_2samp(col_i[st:st + n], col_j[st:st + n])[1])
        o[g] = np.min(ps)
    return o


# --------------------------------------------------------------------------------------
# Stage 3b: binning (phet.py:500-504) and Stage 4: combine (phet.py:510-516)
# --------------------------------------------------------------------------------------
def kbins_uniform(o: np.ndarray, n_bins: int):
    """sklearn KBinsDiscretizer(n_bins, encode='ordinal', strategy='uniform') on one column
    (_discretization.py:325-339 fit, 453-455 transform).  Returns (bins int64 (G,), edges)."""
    o = np.asarray(o, dtype=np.float64)
    col_min, col_max = o.min(), o.max()
    if col_min == col_max:
        return np.zeros(o.shape[0], dtype=np.int64), np.array([-np.inf, np.inf])
    edges = np.linspace(col_min, col_max, n_bins + 1)
    bins = np.searchsorted(edges[1:-1], o, side="right")
    return bins.astype(np.int64), edges


def combine(r: np.ndarray, f: np.ndarray, bins: np.ndarray, feature_weight: np.ndarray):
    """phet.py:502-516.  ``feature_weight`` already normalised to sum 1 (phet.py:124)."""
    G = r.shape[0]
    B = len(feature_weight)
    onehot = np.zeros((G, B), dtype=np.int8)
    for b in range(B):
        onehot[np.where(bins == b)[0], b] = 1
    r = r / r.sum()
    w = feature_weight.dot(onehot.T)
    f = np.multiply(f, w)
    f = f / f.sum()
    res = r + f
    res = np.array(res, copy=True)
    np.nan_to_num(res, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    return np.reshape(res, (G, 1))


# --------------------------------------------------------------------------------------
# Whole path
# --------------------------------------------------------------------------------------
@dataclass
class OracleResult:
    scores: np.ndarray
    m: int
    replace: bool
    idx_i: np.ndarray
    idx_j: np.ndarray
    P: np.ndarray | None
    R: np.ndarray | None
    fsum: np.ndarray
    rsum: np.ndarray
    f: np.ndarray
    r: np.ndarray
    H: np.ndarray
    slices: list
    o: np.ndarray
    bins: np.ndarray
    edges: np.ndarray
    bin_counts: np.ndarray
    extra: dict = field(default_factory=dict)


def fit_predict_oracle(X, y, normalize="zscore", percentiles=(25, 75), num_subsamples=1000,
                       subsampling_size="sqrt", feature_weight=None, rng=np.random,
                       library=True, keep_PR=True, index_sets=None, perms=None) -> OracleResult:
    """Stage-by-stage replay of PHet.fit_predict (binary / iqr path), consuming ``rng`` in the
    reference's order: S x (control choice, case choice), then G x (control perm, case perm).
    ``index_sets=(idx_i, idx_j)`` / ``perms=(perm_i, perm_j)`` bypass the corresponding draws."""
    X = np.asarray(X)
    y = np.asarray(y)
    if feature_weight is None:
        feature_weight = [0.4, 0.3, 0.2, 0.1]
    fw = np.array(feature_weight) / np.sum(feature_weight)
    if normalize == "zscore":
        Xz = zscore_nan_to_num(X)
    elif normalize is None:
        Xz = np.array(X, copy=True)
        np.nan_to_num(Xz, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    else:
        raise ValueError("oracle covers normalize in {'zscore', None}")
    G = X.shape[1]
    S = num_subsamples
    ctrl = np.where(y == 0)[0]
    case = np.where(y == 1)[0]
    m = subsample_size(y, subsampling_size)
    if index_sets is None:
        idx_i, idx_j, replace = draw_index_sets(y, m, S, rng=rng)
    else:
        idx_i, idx_j = index_sets
        replace = bool(m > len(ctrl) or m > len(case))
    P, R = step1_all(Xz, idx_i, idx_j, percentiles, m, library=library)
    fterms = fisher_terms(P)
    fsum = np.sum(fterms, axis=1)
    rsum = np.sum(R, axis=1)
    r = delta_aggregate(R)
    f = fisher_standardise(fsum, S)
    if perms is None:
        perm_i, perm_j = draw_step3_perms(len(ctrl), len(case), G, rng=rng)
    else:
        perm_i, perm_j = perms
    H, slices = step3_h(Xz, ctrl, case, perm_i, perm_j, m)
    o = step3_o_from_h(H, slices)
    bins, edges = kbins_uniform(o, len(fw))
    scores = combine(r, f, bins, fw)
    counts = np.bincount(bins, minlength=len(fw)).astype(np.int64)
    return OracleResult(scores=scores, m=m, replace=replace, idx_i=idx_i, idx_j=idx_j,
                        P=P if keep_PR else None, R=R if keep_PR else None, fsum=fsum, rsum=rsum,
                        f=f, r=r, H=H, slices=slices, o=o, bins=bins, edges=edges,
                        bin_counts=counts, extra={"perm_i": perm_i, "perm_j": perm_j, "Xz": Xz})
