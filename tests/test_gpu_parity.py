"""GPU parity: the CUDA path (through the C ABI / drop-in class) against the committed golden
vectors (outputs of the unmodified reference) and against the oracle on the same inputs.

Integer outputs (index sets, KS statistics h, bins, bin counts, top-k ranking) must be
bit-exact.  Floating-point outputs agree within rtol 1e-5 when the matrix is staged in fp64
(north_star: "rtol 1e-5 in fp64") and rtol 1e-3 when it is staged in fp32.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

RTOL = {"f64": 1e-5, "f32": 1e-3}
NAMES = golden_names()


def _run(g, tag, storage="auto", permutation="reference", capture=True):
    from phet_b200 import PHet
    dt = np.float64 if tag == "f64" else np.float32
    X = np.ascontiguousarray(g["X"].astype(dt))
    np.random.seed(int(g["seed"]))
    est = PHet(**g["params"], storage=storage, permutation=permutation, capture=capture, distributed=False)
    scores = est.fit_predict(X, g["y"].astype(np.int64))
    return est, scores


def _topk(s, k):
    return np.argsort(-s[:, 0], kind="stable")[:k]


@pytest.mark.parametrize("name", NAMES)
def test_scores_match_reference(name):
    g = load_golden(name)
    for tag in g["dtypes"]:
        est, scores = _run(g, tag)
        ref = g[f"scores_{tag}"]
        assert scores.shape == ref.shape and scores.dtype == np.float64
        # draws: identical consumption of np.random
        assert np.array_equal(est.index_sets_[:, 0], g["idx_i"]) and np.array_equal(est.index_sets_[:, 1], g["idx_j"])
        # integer outputs: bit-exact
        assert np.array_equal(est.H_, g[f"H_{tag}"]), "KS integer statistics differ"
        assert np.array_equal(est.bins_.astype(np.int64), g[f"bins_{tag}"].astype(np.int64))
        assert np.array_equal(est.bin_counts_, g[f"bin_counts_{tag}"])
        assert np.array_equal(est.o_, g[f"o_{tag}"]), "o must be bit-exact (table lookup of exact h)"
        if est.o_.min() < est.o_.max():   # edges are formed on the device with numpy's operation order
            assert np.array_equal(est.edges_, np.linspace(est.o_.min(), est.o_.max(), len(est.feature_weight) + 1))
        # floating point
        rtol = RTOL[tag]
        ks = g[f"P_head_{tag}"].shape[1]
        assert np.allclose(est.R_[:, :ks], g[f"R_head_{tag}"], rtol=rtol, atol=1e-5 if tag == "f32" else 1e-12)
        assert np.allclose(est.P_[:, :ks], g[f"P_head_{tag}"], rtol=rtol, atol=1e-300)
        assert np.allclose(est.rsum_, g[f"rsum_{tag}"], rtol=rtol, atol=1e-9)
        assert np.allclose(est.fsum_, g[f"fsum_{tag}"], rtol=rtol, atol=1e-9)
        assert np.allclose(scores, ref, rtol=rtol, atol=1e-12), float(np.max(np.abs(scores - ref) / np.abs(ref).max()))
        k = min(100, scores.shape[0] // 4)
        if tag == "f64":
            assert np.array_equal(_topk(scores, k), _topk(ref, k)), "top-k ranking differs"


@pytest.mark.parametrize("name", ["srbct_s200", "hbecs_like", "cells10k_m70", "unbalanced_2200"])
def test_f32_staging_of_f64_input(name):
    """fp64 input staged as fp32 in HBM/SMEM (the fast layout): rtol 1e-3, exact ranking head."""
    g = load_golden(name)
    est, scores = _run(g, "f64", storage="f32")
    ref = g["scores_f64"]
    assert np.allclose(scores, ref, rtol=1e-3, atol=1e-12)
    assert np.array_equal(est.bin_counts_, g["bin_counts_f64"])
    assert np.array_equal(_topk(scores, 20), _topk(ref, 20))


@pytest.mark.parametrize("name", ["hbecs_like", "balanced_1200", "cells10k_m70"])
def test_device_permutation_mode_against_oracle(name):
    """permutation='device': replay the device's keyed permutations on the host, feed them to the
    oracle, and require identical integer statistics / bins and matching scores."""
    from oracle import phet_oracle as po
    from phet_b200.perm import device_permutation
    g = load_golden(name)
    X = g["X"].astype(np.float64)
    y = g["y"].astype(np.int64)
    np.random.seed(int(g["seed"]))
    from phet_b200 import PHet
    est = PHet(**g["params"], permutation="device", capture=True, distributed=False)
    scores = est.fit_predict(X, y)
    # the class draws its device seed from np.random right after the Step-1 draws
    np.random.seed(int(g["seed"]))
    idx_i, idx_j, _ = po.draw_index_sets(y, int(g["m"]), g["params"]["num_subsamples"])
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    assert seed == est.device_seed_
    ctrl, case = np.where(y == 0)[0], np.where(y == 1)[0]
    G = X.shape[1]
    rows = est.layout_rows_
    assert sorted(rows[:len(ctrl)].tolist()) == ctrl.tolist() and sorted(rows[len(ctrl):].tolist()) == case.tolist()
    perm_i = np.stack([device_permutation(seed, gi, 0, ctrl, rows, 0) for gi in range(G)]).astype(np.int64)
    perm_j = np.stack([device_permutation(seed, gi, 1, case, rows, len(ctrl)) for gi in range(G)]).astype(np.int64)
    assert all(sorted(p.tolist()) == list(range(len(ctrl))) for p in perm_i[:5])
    res = po.fit_predict_oracle(X, y, index_sets=(idx_i, idx_j), perms=(perm_i, perm_j), **g["params"])
    assert np.array_equal(est.H_, res.H)
    assert np.array_equal(est.bin_counts_, res.bin_counts)
    assert np.allclose(scores, res.scores, rtol=1e-5, atol=1e-12)


def test_c_abi_one_call_fit_host():
    """phet_fit_host (the single C entry a non-Python caller would use) == class path."""
    import ctypes as C
    from phet_b200 import _capi, hostmath
    g = load_golden("hbecs_like")
    X = np.ascontiguousarray(g["X"].astype(np.float64))
    y = g["y"].astype(np.int64)
    ctrl = np.where(y == 0)[0].astype(np.int32)
    case = np.where(y == 1)[0].astype(np.int32)
    m = int(g["m"])
    S = g["params"]["num_subsamples"]
    idx = np.ascontiguousarray(np.stack([g["idx_i"], g["idx_j"]], axis=1).astype(np.int32))
    min_c = min(len(ctrl), len(case))
    n_slices, n_last = hostmath.step3_slices(min_c, m)
    p0 = np.ascontiguousarray(g["perm_i"][:, :min_c].astype(np.uint32))
    p1 = np.ascontiguousarray(g["perm_j"][:, :min_c].astype(np.uint32))
    w = np.array([0.4, 0.3, 0.2, 0.1])
    tf, tl = hostmath.ks_exact_table(m), hostmath.ks_exact_table(n_last)
    (j0, k0, g0), (j1, k1, g1) = hostmath.quantile_positions(m, (25, 75))
    prm = _capi.FitParams()
    prm.normalize, prm.storage = _capi.NORM_ZSCORE, _capi.PHET_F64
    prm.step1.test = _capi.TEST_Z
    prm.step1.q = _capi.QuantilePos(j0, k0, j1, k1, g0, g1)
    prm.m, prm.perm_mode, prm.seed, prm.n_bins, prm.n_last = m, _capi.PERM_HOST, 0, 4, n_last
    prm.weights = w.ctypes.data_as(C.POINTER(C.c_double))
    prm.table_full = tf.ctypes.data_as(C.POINTER(C.c_double))
    prm.table_last = tl.ctypes.data_as(C.POINTER(C.c_double))
    ctx = _capi.Context(0)
    scores = np.empty(X.shape[1])
    counts = np.zeros(4, dtype=np.int64)
    rc = ctx.lib.phet_fit_host(ctx.h, _capi._ptr(X), _capi.PHET_F64, X.shape[0], X.shape[1], _capi._ptr(ctrl),
                               len(ctrl), _capi._ptr(case), len(case), _capi._ptr(idx), S, _capi._ptr(p0),
                               _capi._ptr(p1), C.byref(prm), _capi._ptr(scores), _capi._ptr(counts))
    assert rc == 0, ctx.lib.phet_last_error()
    ctx.close()
    assert np.array_equal(counts, g["bin_counts_f64"])
    assert np.allclose(scores[:, None], g["scores_f64"], rtol=1e-5, atol=1e-12)


def test_errors_are_loud():
    from phet_b200 import PHet, _capi
    with pytest.raises(ValueError):
        PHet(feature_weight=[1.0])
    with pytest.raises(Exception):
        PHet().fit_predict(np.zeros((4, 3)), np.zeros(4))
    ctx = _capi.Context(0)
    with pytest.raises(_capi.PhetError):
        ctx.load_matrix(np.zeros((4, 3)))  # classes not set
    ctx.close()


def test_cli_main_srbct(tmp_path):
    """python -m phet_b200.main --methods phet on SRBCT (BASELINE config 0 shape) against the
    reference pipeline's output (float32 load, seed 12345, gamma selection, descending sort)."""
    import os
    from scipy.io import mmwrite
    from scipy.sparse import coo_matrix
    from conftest import GOLDEN_DIR
    from phet_b200.main import main
    g = load_golden("srbct_full")
    z = np.load(os.path.join(GOLDEN_DIR, "cli_srbct.npz"))
    mmwrite(str(tmp_path / "srbct_matrix.mtx"), coo_matrix(g["X"].astype(np.float64)), precision=6)
    with open(tmp_path / "srbct_classes.csv", "w") as fh:
        fh.write("classes\n" + "\n".join(str(int(v)) for v in g["y"]) + "\n")
    with open(tmp_path / "srbct_feature_names.csv", "w") as fh:
        fh.write("features\n" + "\n".join(z["names"].tolist()) + "\n")
    main(["--file-name", "srbct", "--dspath", str(tmp_path), "--rspath", str(tmp_path / "out"), "--methods", "phet",
          "--num-subsamples", str(int(z["S"]))])
    import pandas as pd
    df = pd.read_csv(tmp_path / "out" / "srbct_phet_br_features.csv")
    ref = z["sig_features"].tolist()
    got = df["features"].to_list()
    assert got[:10] == ref[:10]
    jac = len(set(got) & set(ref)) / len(set(got) | set(ref))
    assert jac >= 0.97, jac
    ref_scores = dict(zip(ref, z["sig_scores"]))
    common = [f for f in got if f in ref_scores]
    assert np.allclose(df.set_index("features").loc[common, "score"].to_numpy(), [ref_scores[f] for f in common],
                       rtol=1e-3)


def _adversarial_cases():
    rng = np.random.default_rng(7)
    cases = {}
    # zero-inflated counts: one huge tie group that the percentiles fall into for most subsamples
    X = rng.poisson(0.15, (900, 40)).astype(np.float64)
    cases["zero_inflated"] = (X, 450, dict(num_subsamples=40))
    # a few distinct values only (every bucket is a tie group), plus constant genes
    X = rng.integers(0, 4, (700, 30)).astype(np.float64)
    X[:, :3] = 2.0
    cases["few_levels"] = (X, 300, dict(num_subsamples=30))
    # heavy tails: the sampled window is a poor fit for some genes
    X = rng.standard_cauchy((1000, 30))
    cases["cauchy"] = (X, 500, dict(num_subsamples=30))
    # m not a multiple of 16, wide and narrow percentiles, replace=True (m > class size)
    X = rng.gamma(2.0, 1.0, (600, 24))
    cases["m_37"] = (X, 300, dict(num_subsamples=25, subsampling_size=37, percentiles=(5, 95)))
    cases["m_250"] = (X, 300, dict(num_subsamples=12, subsampling_size=250, percentiles=(40, 60)))
    cases["replace_true"] = (X[:130], 60, dict(num_subsamples=20, subsampling_size=100))
    # bimodal: the two percentiles sit in different modes with a gap between them
    X = np.where(rng.random((800, 20)) < 0.5, rng.normal(-5, 0.1, (800, 20)), rng.normal(5, 0.1, (800, 20)))
    cases["bimodal"] = (X, 400, dict(num_subsamples=30))
    return cases


@pytest.mark.parametrize("name", sorted(_adversarial_cases()))
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_step1_pair_path_equals_sort_path(name, dt, monkeypatch):
    """The bucket-select kernel (k_step1p) must reproduce the sorting kernels exactly: identical
    order statistics (delta-IQR bit for bit) and the same p-values up to fp64 summation order."""
    from phet_b200 import PHet
    X, n0, kw = _adversarial_cases()[name]
    X = np.ascontiguousarray(X.astype(dt))
    y = np.r_[np.zeros(n0, dtype=np.int64), np.ones(X.shape[0] - n0, dtype=np.int64)]
    out = {}
    for impl, nbw in (("auto", "32"), ("auto", "16"), ("warp", "0")):
        monkeypatch.setenv("PHET_STEP1_IMPL", impl)
        monkeypatch.setenv("PHET_STEP1_NBW", nbw)   # histogram width of the select kernel: 127 or 63 buckets
        np.random.seed(99)
        est = PHet(**kw, capture=True, distributed=False, permutation="device")
        est.fit_predict(X, y)
        out[impl + nbw] = (est.R_.copy(), est.P_.copy(), est.step1_geometry_)
    ref = out["warp0"]
    assert ref[2]["path"] != "pair-per-thread"
    for key, buckets in (("auto32", 127), ("auto16", 63)):
        got = out[key]
        assert got[2]["path"] == "pair-per-thread" and got[2]["E"] == buckets
        assert np.array_equal(got[0], ref[0]), "delta-IQR differs between the select and the sort path"
        # float storage of z-scored data sums each 16-element chunk in fp32 (k_step1p FASTMOM): moments to ~1e-7
        assert np.allclose(got[1], ref[1], rtol=1e-11 if dt == np.float64 else 2e-4, atol=1e-300)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("impl", ["scatter", "prune"])
def test_step3_prune_path_matches_reference(name, impl, monkeypatch):
    """Without capture Step 3 runs a bound-and-prune kernel (histogram bounds on h, exact evaluation of
    the slices that can attain the minimum p only) -- the staged gather formulation (k_step3p, the default when the
    class vectors fit in shared memory) or the streaming scatter formulation (k_step3s): o must still be bit-exact, bins and counts identical."""
    g = load_golden(name)
    monkeypatch.setenv("PHET_STEP3_IMPL", impl)
    for tag in g["dtypes"]:
        est, scores = _run(g, tag, capture=False)
        # (gather formulation: a class vector that leaves no room for the histograms falls back to the sorting kernels)
        want = {"scatter": "scatter bound-and-prune", "prune": "bound-and-prune"}[impl]
        assert est.step3_path_ == want or (impl == "prune" and g["X"].shape[0] > 20000), est.step3_path_
        assert np.array_equal(est.o_, g[f"o_{tag}"]), "o must be bit-exact"
        assert np.array_equal(est.bin_counts_, g[f"bin_counts_{tag}"])
        assert np.allclose(scores, g[f"scores_{tag}"], rtol=RTOL[tag], atol=1e-12)


@pytest.mark.parametrize("name", sorted(_adversarial_cases()))
@pytest.mark.parametrize("perm", ["device", "reference"])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_step3_prune_path_equals_sort_path(name, perm, dt, monkeypatch):
    from phet_b200 import PHet
    X, n0, kw = _adversarial_cases()[name]
    X = np.ascontiguousarray(X.astype(dt))
    y = np.r_[np.zeros(n0, dtype=np.int64), np.ones(X.shape[0] - n0, dtype=np.int64)]
    out = {}
    for impl in ("auto", "scatter", "prune", "sort"):
        monkeypatch.setenv("PHET_STEP3_IMPL", impl)
        np.random.seed(5)
        est = PHet(**kw, capture=False, distributed=False, permutation=perm)
        est.fit_predict(X, y)
        out[impl] = (est.o_.copy(), est.step3_path_)
    assert out["auto"][1] == "bound-and-prune" and out["scatter"][1] == "scatter bound-and-prune"
    assert out["prune"][1] == "bound-and-prune" and "bound-and-prune" not in out["sort"][1]
    for impl in ("auto", "scatter", "prune"):
        assert np.array_equal(out[impl][0], out["sort"][0]), impl


def _global_gather_cases():
    """Inputs for the global-gather instantiation of the select kernel (BASELINE config 5's regime:
    m > 255 and/or classes that cannot be staged in shared memory)."""
    rng = np.random.default_rng(21)
    cases = dict(_adversarial_cases())
    X = rng.gamma(2.0, 1.0, (1500, 16))
    cases["m_316"] = (X, 700, dict(num_subsamples=40, subsampling_size=316))
    cases["m_496_p1090"] = (X, 750, dict(num_subsamples=33, subsampling_size=496, percentiles=(10, 90)))
    # tie groups of 256 and more elements in one bucket: the one-byte counters wrap -> full-sort rescue
    X = rng.poisson(0.3, (1600, 12)).astype(np.float64)
    X[:, 0] = 1.0
    cases["m_400_big_ties"] = (X, 800, dict(num_subsamples=35, subsampling_size=400))
    X = rng.integers(0, 3, (1200, 10)).astype(np.float64)
    cases["m_300_three_levels"] = (X, 600, dict(num_subsamples=33, subsampling_size=300, percentiles=(30, 70)))
    return cases


@pytest.mark.parametrize("name", sorted(_global_gather_cases()))
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_step1_global_gather_path_equals_sort_path(name, dt, monkeypatch):
    """k_step1p<BIG> (values gathered from global memory through 32-bit ascending positions, exact word
    sums, wrap detection of the byte counters, 16-bit element numbers) against the sorting kernels."""
    from phet_b200 import PHet
    X, n0, kw = _global_gather_cases()[name]
    X = np.ascontiguousarray(X.astype(dt))
    y = np.r_[np.zeros(n0, dtype=np.int64), np.ones(X.shape[0] - n0, dtype=np.int64)]
    out = {}
    monkeypatch.setenv("PHET_STEP1_FORCE_BIG", "1")
    for impl, nbw in (("auto", "32"), ("auto", "16"), ("warp", "0")):
        monkeypatch.setenv("PHET_STEP1_IMPL", impl)
        monkeypatch.setenv("PHET_STEP1_NBW", nbw)
        np.random.seed(99)
        est = PHet(**kw, capture=True, distributed=False, permutation="device")
        est.fit_predict(X, y)
        out[impl + nbw] = (est.R_.copy(), est.P_.copy(), est.step1_geometry_)
    ref = out["warp0"]
    assert ref[2]["path"] != "pair-per-thread"
    for key, buckets in (("auto32", 127), ("auto16", 63)):
        got = out[key]
        assert got[2]["path"] == "pair-per-thread" and got[2]["E"] == buckets and not got[2]["staged"]
        assert np.array_equal(got[0], ref[0]), "delta-IQR differs between the select and the sort path"
        assert np.allclose(got[1], ref[1], rtol=1e-11 if dt == np.float64 else 2e-4, atol=1e-300)


@pytest.mark.parametrize("name", NAMES)
def test_global_gather_path_matches_reference(name, monkeypatch):
    """Every golden case again with the global-gather instantiation of Step 1 forced."""
    monkeypatch.setenv("PHET_STEP1_FORCE_BIG", "1")
    g = load_golden(name)
    for tag in g["dtypes"]:
        est, scores = _run(g, tag)
        assert est.step1_geometry_["path"] == "pair-per-thread" and not est.step1_geometry_["staged"]
        ks = g[f"P_head_{tag}"].shape[1]
        rtol = RTOL[tag]
        assert np.allclose(est.R_[:, :ks], g[f"R_head_{tag}"], rtol=rtol, atol=1e-5 if tag == "f32" else 1e-12)
        assert np.allclose(est.P_[:, :ks], g[f"P_head_{tag}"], rtol=rtol, atol=1e-300)
        assert np.array_equal(est.bin_counts_, g[f"bin_counts_{tag}"])
        assert np.allclose(scores, g[f"scores_{tag}"], rtol=rtol, atol=1e-12)


@pytest.mark.parametrize("impl", ["auto", "warp"])
def test_big_class_matches_oracle(impl, monkeypatch):
    """Classes above 65,536 cells (BASELINE config 5's regime) cannot use the 16-bit index table and do
    not fit shared memory.  140,000 cells -> m = 264, 266 KS slices.  auto: the global-gather select kernel;
    warp: the warp-cooperative sorters gathering from L2 (the fallback).  Both against the oracle."""
    from oracle import phet_oracle as po
    from phet_b200 import PHet
    monkeypatch.setenv("PHET_STEP1_IMPL", impl)
    rng = np.random.default_rng(11)
    N, G, S = 140000, 6, 5
    y = np.r_[np.zeros(N // 2, dtype=np.int64), np.ones(N - N // 2, dtype=np.int64)]
    X = rng.gamma(2.0, 1.0, (N, G))
    X[y == 1, 0] += 0.05
    np.random.seed(3)
    est = PHet(num_subsamples=S, capture=True, distributed=False)
    scores = est.fit_predict(X, y)
    assert not est.step1_geometry_["staged"]
    assert est.step1_geometry_["path"] == ("pair-per-thread" if impl == "auto" else "full-warp")
    np.random.seed(3)
    ref = po.fit_predict_oracle(X, y, num_subsamples=S)
    assert np.array_equal(est.H_, ref.H)
    assert np.array_equal(est.bin_counts_, ref.bin_counts)
    assert np.allclose(est.R_, ref.R, rtol=1e-5, atol=1e-12) and np.allclose(est.P_, ref.P, rtol=1e-5, atol=1e-300)
    assert np.allclose(scores, ref.scores, rtol=1e-5, atol=1e-12)
