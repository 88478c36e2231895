"""GPU parity of the sibling IQR scorers (phet_b200.siblings -> phet_class_stats) against outputs of the
unmodified reference classes (tests/golden/siblings_*.npz).

Bar: class_diff (integer column) bit-exact; per-class IQRs bit-exact when the matrix is not normalised
(exact order statistics + scipy's interpolation order in X's dtype); everything else rtol 1e-5 for float64
input and 1e-3 for float32 input (the reference z-scores and averages float32 data in float32)."""
import numpy as np
import pytest

from test_siblings_cpu import CASES, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_class_stats_and_scores(name):
    from phet_b200 import HIQR, DeltaIQRMean
    z, norm, rng_ = load(name)
    X, y = z["X"], z["y"]
    f32 = X.dtype == np.float32
    rtol, atol = (1e-3, 1e-5) if f32 else (1e-5, 1e-11)
    for pc in (False, True):
        est = HIQR(per_condition=pc, normalize=norm, iqr_range=rng_)
        got = est.fit_predict(X.copy(), y)
        ref = z["hiqr_pc%d" % pc]
        assert got.shape == ref.shape and got.dtype == np.float64 and est.launches_ > 0
        assert np.allclose(got, ref, rtol=rtol, atol=atol)
    st = est.class_stats_
    if norm is None:   # no normalisation: the device sees the reference's values, so order statistics AND lerp are exact
        assert np.array_equal(st[0], z["st_iqr0"]) and np.array_equal(st[1], z["st_iqr1"]) and np.array_equal(st[2], z["st_iqr_all"])
    for k, key in enumerate(("iqr0", "iqr1", "iqr_all", "mean0", "mean1", "ss0", "ss1")):
        ref = np.nan_to_num(z["st_" + key], nan=0.0)   # constant genes: NaN in the reference, 0 here (same final columns)
        assert np.allclose(st[k], ref, rtol=rtol, atol=1e-4 if f32 else 1e-9), key
    for dm in (False, True):
        got = DeltaIQRMean(calculate_deltamean=dm, normalize=norm, iqr_range=rng_).fit_predict(X.copy(), y)
        ref = z["dim_dm%d" % dm]
        assert got.shape == ref.shape
        assert np.array_equal(got[:, 4], ref[:, 4]), "class_diff must be bit-exact"
        assert np.allclose(got[:, :4], ref[:, :4], rtol=rtol, atol=atol), np.max(np.abs(got[:, :4] - ref[:, :4]), axis=0)
        k = 10
        assert np.array_equal(np.argsort(-got[:, 3], kind="stable")[:k], np.argsort(-ref[:, 3], kind="stable")[:k])


def test_large_vectors_against_numpy():
    """40,000-cell classes (radix select over vectors far beyond the sort kernels' reach), float32 and
    float64, with heavy ties; order statistics checked against numpy's sort."""
    from phet_b200 import _capi, hostmath
    rng = np.random.default_rng(5)
    N, G = 90000, 12
    y = (rng.random(N) < 0.45).astype(np.int64)
    for dt in (np.float32, np.float64):
        X = rng.gamma(2.0, 1.0, (N, G)).astype(dt)
        X[:, 0] = np.round(X[:, 0])          # ties
        X[:, 1] = 7.0                        # constant
        X[:, 2] *= -1.0                      # negative values (key transform)
        i0, i1 = np.where(y == 0)[0], np.where(y == 1)[0]
        q3 = []
        for n in (len(i0), len(i1), N):
            (j0, k0, g0), (j1, k1, g1) = hostmath.quantile_positions(n, (25, 75), dtype=dt)
            q3.append(_capi.QuantilePos(j0, k0, j1, k1, g0, g1))
        ctx = _capi.Context(0)
        try:
            ctx.set_classes(i0, i1)
            ctx.load_matrix(X, normalize=_capi.NORM_NONE)
            st = ctx.class_stats(q3)
        finally:
            ctx.close()
        from scipy.stats import iqr
        for k, rows in enumerate((i0, i1, np.arange(N))):
            ref = iqr(X[rows], axis=0, rng=(25, 75), scale=1.0)
            assert np.array_equal(st[k], ref.astype(np.float64)), (dt, k)
        assert np.allclose(st[3], X[i0].astype(np.float64).mean(0), rtol=1e-12)
        assert np.allclose(st[6], ((X[i1].astype(np.float64) - X[i1].astype(np.float64).mean(0)) ** 2).sum(0), rtol=1e-9, atol=1e-9)


def test_scope_errors_are_loud():
    from phet_b200 import HIQR, DeltaIQRMean
    X = np.random.default_rng(0).random((30, 4))
    with pytest.raises(NotImplementedError):
        HIQR(normalize="robust").fit_predict(X, np.r_[np.zeros(15), np.ones(15)].astype(int))
    with pytest.raises(NotImplementedError):
        DeltaIQRMean().fit_predict(X, np.arange(30) % 3)
