"""Host-side scalar arithmetic of the PHet path: everything that is O(1) or O(m) per fit and
therefore stays on the CPU, written to reproduce the third-party functions the reference calls.

* ``subsample_size``       -- /root/reference/src/model/phet.py:126-161, 309-332
* ``quantile_positions``   -- scipy.stats.iqr -> stats.quantile(method='linear') index arithmetic
                              (scipy/stats/_quantile.py ``_quantile_hf``), reached from phet.py:413-414
* ``ks_exact_table``       -- scipy.stats.ks_2samp two-sided exact p-value for n1 == n2 as a function
                              of the integer statistic (scipy/stats/_stats_py.py
                              ``_compute_prob_outside_square`` / ``_attempt_exact_2kssamp`` /
                              ``_ks_2samp_prob``), reached from phet.py:493
* ``uniform_bin_edges``    -- sklearn KBinsDiscretizer(strategy='uniform') edges, phet.py:500-501
"""
from __future__ import annotations

import math

import numpy as np


def optimum_sample_size(control_size: int, case_size: int, alpha: float = 0.05,
                        margin_error: float = 0.1) -> int:
    from scipy.stats import norm
    total_samples = control_size + case_size
    p = control_size / total_samples
    q = 1 - p
    z = norm.ppf(1 - alpha / 2)
    numerator = (z ** 2 * p * q) / (margin_error ** 2)
    denominator = 1 + (numerator / total_samples)
    return int(np.ceil(numerator / denominator))


def subsample_size(n0: int, n1: int, subsampling_size) -> int:
    if isinstance(subsampling_size, str):
        if subsampling_size == "optimum":
            return int(optimum_sample_size(n0, n1, alpha=0.05, margin_error=0.1))
        if subsampling_size == "sqrt":
            min_c = min(n0, n1)
            if min_c <= 2 or int(np.sqrt(min_c)) < 8:
                return int(min_c)
            return int(np.sqrt(min_c))
        raise ValueError("subsampling_size must be 'optimum', 'sqrt' or an int")
    return int(subsampling_size)


def quantile_positions(m: int, percentiles, dtype=np.float64):
    """((j_lo, k_lo, g_lo), (j_hi, k_hi, g_hi)); arithmetic in ``dtype`` like scipy does in X's dtype."""
    lo, hi = percentiles
    if np.isnan([lo, hi]).any():
        raise ValueError("`rng` must not contain NaNs.")
    rng = (lo / 100, hi / 100) if lo < hi else (hi / 100, lo / 100)
    if rng[0] < 0 or rng[1] > 1:
        raise ValueError("Elements of `rng` must be in the range [0, 100].")
    out = []
    n = dtype(m)
    for p in rng:
        p = dtype(p)
        jg = p * n + (dtype(1) - p)
        jp1 = np.floor(jg)
        j = jp1 - 1
        g = jg - jp1
        if j < 0:
            g = dtype(0)
        j = min(max(j, 0.0), n - 1)
        jp1 = min(max(jp1, 0.0), n - 1)
        out.append((int(j), int(jp1), float(g)))
    return tuple(out)


def ln_beta_half(df: float) -> float:
    """ln B(df/2, 1/2)."""
    a = 0.5 * df
    return math.lgamma(a) + math.lgamma(0.5) - math.lgamma(a + 0.5)


def _prob_outside_square(n: int, h: int) -> float:
    P = 0.0
    k = int(np.floor(n / h))
    while k >= 0:
        p1 = 1.0
        for j in range(h):
            p1 = (n - k * h - j) * p1 / (n + k * h + j + 1)
        P = p1 * (1.0 - P)
        k -= 1
    return 2 * P


_KS_CACHE: dict[int, np.ndarray] = {}


def ks_exact_table(n: int) -> np.ndarray:
    """table[h] = ks_2samp(a, b).pvalue for len(a) == len(b) == n and n*D == h (two-sided, 'auto')."""
    if n in _KS_CACHE:
        return _KS_CACHE[n]
    if n > 10000:
        # ks_2samp(method="auto") switches to the asymptotic formula above 10000 cells (_stats_py.py:8146-8180):
        # kstwo.sf(d, round(n1 n2 / (n1 + n2))) -- only a multi-class fit's long last slice gets here.  scipy feeds
        # it d = |ca/n - cb/n| (two rounded quotients), the table h/n: equal up to the last ulp of d, i.e. ~1e-14
        # relative in p (the exact branch below is a function of the integer h and is bit-identical)
        from scipy.stats import distributions
        d = np.arange(n + 1, dtype=np.float64) / n
        tab = np.clip(distributions.kstwo.sf(d, np.round(float(n) * float(n) / (float(n) + float(n)))), 0, 1)
        _KS_CACHE[n] = tab
        return tab
    tab = np.empty(n + 1, dtype=np.float64)
    tab[0] = 1.0
    for h in range(1, n + 1):
        prob = _prob_outside_square(n, h)
        if not (0 <= prob <= 1):
            from scipy.stats import distributions
            en = float(n) * float(n) / (float(n) + float(n))
            prob = distributions.kstwo.sf(h * 1.0 / n, np.round(en))
        tab[h] = np.clip(prob, 0, 1)
    _KS_CACHE[n] = tab
    return tab


def uniform_bin_edges(o_min: float, o_max: float, n_bins: int) -> np.ndarray:
    if o_min == o_max:
        return np.array([-np.inf, np.inf])
    return np.linspace(o_min, o_max, n_bins + 1)


def step3_slices(min_c: int, m: int):
    n_slices = (min_c + m - 1) // m
    n_last = min_c - (n_slices - 1) * m
    return n_slices, n_last
