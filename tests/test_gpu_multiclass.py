"""GPU: multi-class branch through the C ABI against the fixtures produced by the unmodified
reference (tests/golden/make_golden_multiclass.py).  Integer outputs (KS statistics, bins, bin counts, ranking)
must be identical; floating-point outputs agree within rtol 1e-5 for float64 input and 1e-3 for float32 input."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[len("multi_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "multi_*.npz")))


def _fit(z, **kw):
    from phet_b200 import PHet
    est = PHet(num_subsamples=int(z["S"]), percentiles=tuple(z["percentiles"]), multiconditions_strategy=str(z["strategy"]),
               multiconditions_aggregation=str(z["aggregation"]), group_test=str(z["group_test"]), capture=True, device=0, **kw)
    np.random.seed(12345)
    scores = est.fit_predict(z["X"], z["y"].astype(np.int64))
    return est, scores


@pytest.mark.parametrize("name", CASES)
def test_multiclass_matches_reference(name):
    z = np.load(os.path.join(GOLDEN_DIR, "multi_%s.npz" % name))
    f32 = z["X"].dtype == np.float32
    rtol = 1e-3 if f32 else 1e-5
    est, scores = _fit(z)
    assert est.m_ == int(z["m"])
    # Step 1a: ANOVA p-values and Fisher sums
    # (a statistic that rounds to a tiny negative F is a domain error -> NaN -> 0 in cephes fdtrc and p = 1 when it
    # rounds to a tiny positive one: both contribute 0 to Fisher's sum, so the comparison is on the Fisher terms)
    def terms(P):
        with np.errstate(divide="ignore"):
            t = -2 * np.log(P)
        t[t == np.inf] = 0
        return t
    P = z["P"]
    both = (P > 1e-290) & (est.P_ > 0) & (P < 0.999)
    assert both.mean() > 0.4
    assert np.allclose(est.P_[both], P[both], rtol=5e-3 if f32 else 1e-9, atol=0)
    assert np.allclose(terms(est.P_), terms(P), rtol=5e-3 if f32 else 1e-9, atol=1e-6)
    assert np.allclose(est.fsum_, z["fsum"], rtol=rtol, atol=1e-9)
    # Step 1b: delta-IQR sums per combination
    assert np.allclose(est.rsum_combo_, z["rsum"], rtol=1e-5 if f32 else 1e-12, atol=1e-12)
    # Step 3: integer KS statistics per combination and slice, and the p-values looked up from them
    for c in range(z["H"].shape[1]):
        assert np.array_equal(est.H_combo_[c], z["H"][:, c, :]), c
    assert np.array_equal(est.o_combo_, z["o_combo"])
    assert np.array_equal(est.o_, z["o"])
    assert np.array_equal(est.bins_, z["bins"])
    assert np.array_equal(est.bin_counts_, z["bin_counts"])
    assert np.allclose(est.edges_, z["edges"], rtol=0, atol=0)
    assert np.allclose(est.r_, z["r"], rtol=1e-5 if f32 else 1e-12, atol=1e-15)
    assert np.allclose(est.f_, z["f"], rtol=rtol, atol=1e-9)
    assert scores.shape == z["scores_ref"].shape
    assert np.allclose(scores, z["scores_ref"], rtol=rtol, atol=1e-12)
    if not f32:
        order = np.argsort(-z["scores_ref"][:, 0], kind="stable")
        assert np.array_equal(est.ranking_[:10], order[:10])


def test_multiclass_without_capture_and_device_permutations():
    """The production configuration: no per-slice capture (the pruning kernel may run), device-side permutations.
    The permutation stream differs from NumPy's, so only the permutation-independent parts are compared."""
    z = np.load(os.path.join(GOLDEN_DIR, "multi_k3_sqrt.npz"))
    from phet_b200 import PHet
    np.random.seed(12345)
    est = PHet(num_subsamples=int(z["S"]), device=0)
    ref_scores = est.fit_predict(z["X"], z["y"].astype(np.int64))
    assert np.allclose(ref_scores, z["scores_ref"], rtol=1e-5, atol=1e-12)
    np.random.seed(12345)
    dev = PHet(num_subsamples=int(z["S"]), device=0, permutation="device")
    s2 = dev.fit_predict(z["X"], z["y"].astype(np.int64))
    assert s2.shape == ref_scores.shape and np.isfinite(s2).all() and abs(s2.sum() - 2.0) < 1e-9
    assert dev.bin_counts_.sum() == z["X"].shape[1]


def test_anova_constant_groups_and_large_statistics():
    """f_oneway's constant-input rules (all groups constant -> p = 0; everything identical -> NaN -> 0) and p-values
    far down the tail, straight through phet_anova against scipy."""
    from scipy import stats
    import math
    from phet_b200 import _capi
    rng = np.random.default_rng(0)
    K, m, S, G = 3, 12, 16, 8
    N = 90
    y = np.repeat(np.arange(K), N // K)
    X = rng.normal(size=(N, G))
    X[:, 0] = 2.0                                  # identical everywhere
    X[:, 1] = y * 1.5                              # constant inside each class
    X[:, 2] += y * 8.0                             # very large F
    X[:, 3] += y * 40.0                            # p far below 1e-30
    idx = np.empty((S, K, m), dtype=np.int32)
    for s in range(S):
        for c in range(K):
            idx[s, c] = rng.choice(np.where(y == c)[0], m, replace=False)
    ctx = _capi.Context(0)
    try:
        ctx.set_classes(np.where(y == 0)[0], np.where(y != 0)[0])
        ctx.load_matrix(X, _capi.NORM_NONE)
        a, b = 0.5 * (K * m - K), 0.5 * (K - 1)
        ctx.anova(idx, math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b), 0.0, capture=True)
        P = ctx.read(_capi.BUF_P, np.float64, (G, S))
        fsum = ctx.read(_capi.BUF_ACC, np.float64, (2, G))[0]
    finally:
        ctx.close()
    want = np.zeros((G, S))
    for s in range(S):
        with np.errstate(all="ignore"):
            _, p = stats.f_oneway(*[X[idx[s, c]] for c in range(K)])
        want[:, s] = np.nan_to_num(p, nan=0.0, posinf=0.0, neginf=0.0)
    assert np.array_equal(P[:2], np.zeros((2, S))) and np.array_equal(want[:2], np.zeros((2, S)))
    assert np.allclose(P, want, rtol=1e-9, atol=0)
    with np.errstate(divide="ignore"):
        t = -2 * np.log(want)
    t[t == np.inf] = 0
    assert np.allclose(fsum, t.sum(1), rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("N", [50000, 120000])   # 200 KB vector staged in shared memory / gathered from L2
def test_group_tests_on_long_cell_vectors(N):
    """phet_anova / phet_kruskal at BASELINE-sized cell counts against scipy on every gene (float32 input: rtol 1e-3)."""
    from scipy import stats
    import math
    from phet_b200 import _capi
    rng = np.random.default_rng(N)
    K, G, S = 3, 48, 40
    y = (np.arange(N) * K // N).astype(np.int64)
    X = rng.gamma(2.0, 1.0, (N, G)).astype(np.float32)
    X[y == 2, :8] += 0.4
    m = int(np.sqrt(min(np.bincount(y))))
    idx = np.empty((S, K, m), dtype=np.int32)
    for s in range(S):
        for c in range(K):
            idx[s, c] = rng.choice(np.where(y == c)[0], m, replace=False)
    ctx = _capi.Context(0)
    try:
        ctx.set_classes(np.where(y == 0)[0], np.where(y != 0)[0])
        ctx.load_matrix(X, _capi.NORM_ZSCORE)
        a, b = 0.5 * (K * m - K), 0.5 * (K - 1)
        ctx.anova(idx, math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b), 2.0 ** -150, capture=True)
        Pa = ctx.read(_capi.BUF_P, np.float64, (G, S))
        ctx.kruskal(idx, 2.0 ** -150, capture=True)
        Pk = ctx.read(_capi.BUF_P, np.float64, (G, S))
    finally:
        ctx.close()
    Xz = stats.zscore(X.astype(np.float64), axis=0)
    for s in range(S):
        groups = [Xz[idx[s, c]] for c in range(K)]
        assert np.allclose(Pa[:, s], stats.f_oneway(*groups)[1], rtol=1e-3, atol=1e-300)
        assert np.allclose(Pk[:, s], stats.kruskal(*groups)[1], rtol=1e-3, atol=1e-300)


@pytest.mark.parametrize("case", ["continuous_m57", "counts_m40", "ties_m20_k5", "f64_m130", "full_strip_m16_k4", "full_strip_m64_k2"])
def test_kruskal_sort_kernel_equals_counting_kernel(case, monkeypatch):
    """k_kruskal_sort (warp sort + two binary searches per element) forms the same exact integers as the counting kernel
    k_kruskal (#{x < v}, #{x <= v}): the p-values must be bit-identical; and they agree with scipy.stats.kruskal."""
    from scipy import stats
    from phet_b200 import _capi
    rng = np.random.default_rng(11)
    K, m, dt, gen = {"continuous_m57": (3, 57, np.float32, lambda s: rng.gamma(2.0, 1.0, s)),
                     "counts_m40": (3, 40, np.float32, lambda s: rng.poisson(0.8, s).astype(np.float64)),
                     "ties_m20_k5": (5, 20, np.float32, lambda s: rng.integers(0, 4, s).astype(np.float64)),
                     "f64_m130": (2, 130, np.float64, lambda s: rng.normal(0, 1, s)),
                     # K m a power of two: the sorted strip has no padding at all
                     "full_strip_m16_k4": (4, 16, np.float32, lambda s: rng.poisson(1.5, s).astype(np.float64)),
                     "full_strip_m64_k2": (2, 64, np.float64, lambda s: rng.integers(0, 6, s).astype(np.float64))}[case]
    N, G, S = 1500, 24, 37
    y = (np.arange(N) * K // N).astype(np.int64)
    X = np.ascontiguousarray(gen((N, G)).astype(dt))
    X[:, 0] = 1.0   # a constant gene: one tie group
    idx = np.empty((S, K, m), dtype=np.int32)
    for s in range(S):
        for c in range(K):
            idx[s, c] = rng.choice(np.where(y == c)[0], m, replace=False)
    P = {}
    for impl in ("sort", "count"):
        monkeypatch.setenv("PHET_KRUSKAL_IMPL", impl)
        ctx = _capi.Context(0)
        try:
            ctx.set_classes(np.where(y == 0)[0], np.where(y != 0)[0])
            ctx.load_matrix(X, _capi.NORM_NONE)
            ctx.kruskal(idx, 0.0, capture=True)
            P[impl] = ctx.read(_capi.BUF_P, np.float64, (G, S))
        finally:
            ctx.close()
    assert np.array_equal(P["sort"], P["count"])
    for s in range(0, S, 6):
        for g in range(1, G, 5):
            want = stats.kruskal(*[X[idx[s, c], g].astype(np.float64) for c in range(K)])[1]
            assert np.isclose(P["sort"][g, s], want, rtol=1e-9, atol=1e-300)
