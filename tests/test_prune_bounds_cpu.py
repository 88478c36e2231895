"""CPU: the inequality the bound-and-prune Step-3 kernel (csrc/step3p_impl.cuh) relies on.

For two samples a, b of equal length bucketed by ANY monotone map with counts ha_k, hb_k and
D_k = sum_{i<=k}(ha_i - hb_i), the integer KS statistic h = max_v |#{a<=v} - #{b<=v}| satisfies
    L = max_k |D_k|  <=  h  <=  U = max_k max(D_{k-1} + ha_k, hb_k - D_{k-1}).
Also checks the table property the pruning needs: ks_exact_table(n) is non-increasing from a small h0 on
and table[h0] is not above any earlier entry (the exact formula has 1-ulp noise at h = 1, 2 for some n)."""
import numpy as np
import pytest

from phet_b200 import hostmath


def ks_h(a, b):
    """Integer two-sample KS statistic for equal sizes (scipy's d * n), ties counted at the same point."""
    v = np.unique(np.concatenate([a, b]))
    ca = np.searchsorted(np.sort(a), v, side="right")
    cb = np.searchsorted(np.sort(b), v, side="right")
    return int(np.max(np.abs(ca - cb)))


def bounds(a, b, edges):
    ha = np.bincount(np.searchsorted(edges, a, side="right"), minlength=len(edges) + 1)
    hb = np.bincount(np.searchsorted(edges, b, side="right"), minlength=len(edges) + 1)
    D = np.cumsum(ha - hb)
    Dprev = np.concatenate([[0], D[:-1]])
    return int(np.max(np.abs(D))), int(np.max(np.maximum(Dprev + ha, hb - Dprev)))


@pytest.mark.parametrize("seed", range(8))
def test_histogram_bounds_bracket_the_ks_statistic(seed):
    rng = np.random.default_rng(seed)
    for n in (1, 7, 36, 158, 255):
        for kind in range(4):
            if kind == 0:
                a, b = rng.normal(size=n), rng.normal(0.3, 1.2, size=n)
            elif kind == 1:   # heavy ties
                a, b = rng.integers(0, 4, n).astype(float), rng.integers(0, 5, n).astype(float)
            elif kind == 2:   # zero inflated
                a, b = rng.poisson(0.2, n).astype(float), rng.poisson(0.4, n).astype(float)
            else:             # disjoint supports
                a, b = rng.uniform(0, 1, n), rng.uniform(2, 3, n)
            lo, hi = np.quantile(np.concatenate([a, b]), [0.005, 0.995])
            for nb in (2, 16, 127):
                edges = np.linspace(lo, hi, nb - 1) if hi > lo else np.array([lo])
                L, U = bounds(a, b, edges)
                h = ks_h(a, b)
                assert L <= h <= U, (n, kind, nb, L, h, U)


@pytest.mark.parametrize("n", [1, 2, 5, 29, 36, 45, 70, 122, 144, 158, 200, 255])
def test_ks_table_is_non_increasing_after_a_small_h0(n):
    t = hostmath.ks_exact_table(n)
    h0 = 0
    for h in range(n):
        if t[h + 1] > t[h]:
            h0 = h + 1
    floor0 = min(t[:h0]) if h0 else np.inf
    while h0 <= n and t[h0] > floor0:
        h0 += 1
    assert h0 <= 8, "the prune threshold h0 must stay tiny (below it every slice is evaluated exactly)"
    assert np.all(np.diff(t[h0:]) <= 0)


def test_ks_table_above_10000_cells_follows_scipys_asymptotic_branch():
    """ks_2samp(method='auto') is asymptotic above 10000 cells; the table reproduces it to ~1e-14 (a multi-class
    fit's long last slice is the only caller that can get there)."""
    import numpy as np
    from scipy import stats
    from phet_b200 import hostmath
    rng = np.random.default_rng(0)
    n = 10007
    tab = hostmath.ks_exact_table(n)
    assert tab.shape == (n + 1,) and tab[0] == 1.0 and np.all(np.diff(tab) <= 0)
    for shift in (0.0, 0.01, 0.03, 0.1):
        a, b = rng.normal(size=n), rng.normal(size=n) + shift
        r = stats.ks_2samp(a, b)
        h = int(round(r.statistic * n))
        assert np.isclose(tab[h], r.pvalue, rtol=1e-12, atol=0)
