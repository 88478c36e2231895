"""CPU: the oracle (oracle/phet_oracle.py) against the committed golden vectors, which are
outputs of the unmodified reference (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import phet_oracle as po

FAST = [n for n in golden_names() if n not in ("srbct_full",)]


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_reference_scores(name):
    g = load_golden(name)
    for tag in g["dtypes"]:
        dt = np.float64 if tag == "f64" else np.float32
        X = np.ascontiguousarray(g["X"].astype(dt))
        np.random.seed(int(g["seed"]))
        res = po.fit_predict_oracle(X, g["y"].astype(np.int64), **g["params"])
        assert np.array_equal(res.idx_i, g["idx_i"]) and np.array_equal(res.idx_j, g["idx_j"])
        assert np.array_equal(res.extra["perm_i"], g["perm_i"])
        assert np.array_equal(res.H, g[f"H_{tag}"])
        assert np.array_equal(res.bin_counts, g[f"bin_counts_{tag}"])
        assert np.max(np.abs(res.scores - g[f"scores_{tag}"])) == 0.0
        assert res.scores.shape == (X.shape[1], 1)


def test_oracle_srbct_full_from_stored_draws():
    """BASELINE config 1 (83x2308, S=1000): stored index sets and perms -> identical scores."""
    g = load_golden("srbct_full")
    res = po.fit_predict_oracle(g["X"].astype(np.float64), g["y"].astype(np.int64),
                                index_sets=(g["idx_i"].astype(np.int64), g["idx_j"].astype(np.int64)),
                                perms=(g["perm_i"].astype(np.int64), g["perm_j"].astype(np.int64)),
                                keep_PR=False, **g["params"])
    assert np.max(np.abs(res.scores - g["scores_f64"])) == 0.0
    assert np.array_equal(np.argsort(-res.scores[:, 0], kind="stable")[:100],
                          np.argsort(-g["scores_f64"][:, 0], kind="stable")[:100])


@pytest.mark.parametrize("name", ["srbct_s200", "hbecs_like", "cells10k_m70"])
def test_plain_restatement_matches_library(name):
    """The formula-level restatements (the arithmetic the CUDA kernels implement) agree with
    the scipy calls the reference makes: rtol 1e-12 in fp64."""
    g = load_golden(name)
    X = g["X"].astype(np.float64)
    Xz = po.zscore_nan_to_num(X)
    assert np.allclose(po.zscore_plain(X), Xz, rtol=1e-12, atol=1e-14)
    m = int(g["m"])
    pct = g["params"]["percentiles"]
    for s in range(3):
        p_lib, d_lib = po.step1_subsample(Xz, g["idx_i"][s], g["idx_j"][s], pct, m, library=True)
        p_pl, d_pl = po.step1_subsample(Xz, g["idx_i"][s], g["idx_j"][s], pct, m, library=False)
        assert np.allclose(p_pl, p_lib, rtol=1e-12, atol=1e-300)
        assert np.allclose(d_pl, d_lib, rtol=1e-12, atol=1e-15)
        assert np.array_equal(p_lib, g["P_head_f64"][:, s])
        assert np.array_equal(d_lib, g["R_head_f64"][:, s])


@pytest.mark.parametrize("n", [1, 2, 8, 29, 45, 70, 158])
def test_ks_table_matches_scipy(n):
    """ks_table(n)[h] == scipy.stats.ks_2samp p-value for samples realising statistic h/n."""
    from scipy import stats
    tab = po.ks_table(n)
    assert tab[0] == 1.0 and np.all(np.diff(tab) <= 0)
    for h in sorted({0, 1, 2, n // 3, n // 2, n - 1, n}):
        if h < 0 or h > n:
            continue
        a = np.arange(n, dtype=np.float64)              # b: h values below all of a, rest interleaved above
        b = np.concatenate([np.full(h, -1.0), np.arange(n - h) + 0.5 + h - 1]) if h > 0 else a + 0.5
        hh = po.ks_h_equal_n(a, b)
        p = stats.ks_2samp(a, b)[1]
        assert tab[hh] == p


def test_permutation_replay_equivalence():
    """np.random.permutation(values) == values[np.random.permutation(len)] with equal state advance."""
    v = np.random.default_rng(0).normal(size=57)
    for dt in (np.float64, np.float32):
        np.random.seed(7)
        a = np.random.permutation(v.astype(dt))
        s1 = np.random.get_state()[1].copy()
        np.random.seed(7)
        b = v.astype(dt)[np.random.permutation(len(v))]
        s2 = np.random.get_state()[1].copy()
        assert np.array_equal(a, b) and np.array_equal(s1, s2)


def test_quantile_positions():
    assert po.quantile_positions(29, (25, 75)) == ((7, 8, 0.0), (21, 22, 0.0))
    (j0, k0, g0), (j1, k1, g1) = po.quantile_positions(70, (25, 75))
    assert (j0, k0, j1, k1) == (17, 18, 51, 52) and abs(g0 - 0.25) < 1e-12 and abs(g1 - 0.75) < 1e-12
    (j0, k0, g0), (j1, k1, g1) = po.quantile_positions(158, (25, 75))
    assert (j0, k0, j1, k1) == (39, 40, 117, 118)
