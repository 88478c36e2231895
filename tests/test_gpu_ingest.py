"""GPU: the rows of SURVEY.md 8(f) either side of the fit, through the C ABI.

* phet_mtx_stage (input path) against the host restatement of /root/reference/src/train.py:77-118 in
  oracle/ingest_oracle.py (load_dataset + filter_data: scipy mmread, float32, nan_to_num, cell filter, CPM gene
  filter) -- the staged matrix and both index lists must be identical, bit for bit;
* phet_rank (ordering of utils.py:107-136) against pandas' descending sort: identical where scores are distinct,
  ascending gene order inside tie groups (pandas kind="stable").
"""
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore::DeprecationWarning")]


def _write_case(tmp_path, name, rng):
    path = str(tmp_path / (name + "_matrix.mtx"))
    if name == "counts":          # zero-inflated counts, some cells below the sum-5 cut, rare genes
        N, G = 400, 700
        X = rng.poisson(0.05, (N, G)).astype(np.float64) * rng.integers(1, 20, (N, G))
        X[:25] *= (rng.random((25, G)) < 0.01)
        X[:, :40] *= (rng.random((N, 40)) < 0.02)
    elif name == "wide":          # several summation leaves per row and non-multiple-of-8 remainders
        N, G = 300, 5003
        X = rng.gamma(0.3, 2.0, (N, G)) * (rng.random((N, G)) < 0.1)
        X[:10] *= 1e-3
    elif name == "border":        # row sums close to 5 and CPM values close to 1
        N, G = 64, 1000
        X = np.zeros((N, G))
        for i in range(N):
            cols = rng.choice(G, 50, replace=False)
            X[i, cols] = rng.random(50)
            X[i, cols] *= (5.0 + (i - 32) * 1e-7) / X[i, cols].sum()
        X[:, 0] = 5e-6
        X[:, 1] = 5.0000005e-6
    elif name == "tiny":          # fewer than minimum_samples cells -> threshold becomes cells // 2
        N, G = 4, 9
        X = rng.poisson(3.0, (N, G)).astype(np.float64)
    else:
        raise KeyError(name)
    scipy.io.mmwrite(path, sp.coo_matrix(X), precision=9)
    y = (np.arange(N) >= N // 2).astype(np.int64)
    with open(tmp_path / (name + "_classes.csv"), "w") as fh:
        fh.write("classes\n" + "\n".join(str(int(v)) for v in y) + "\n")
    with open(tmp_path / (name + "_feature_names.csv"), "w") as fh:
        fh.write("features\n" + "\n".join("g%d" % j for j in range(G)) + "\n")
    return path


@pytest.mark.parametrize("do_filter", [True, False])
@pytest.mark.parametrize("name", ["counts", "wide", "border", "tiny"])
def test_stage_matches_host_filter(tmp_path, name, do_filter):
    from phet_b200 import load_mtx
    from oracle.ingest_oracle import filter_data, load_dataset
    rng = np.random.default_rng(11)
    path = _write_case(tmp_path, name, rng)
    X, y, names = load_dataset(str(tmp_path), name)
    ids = np.arange(X.shape[0])
    if do_filter:
        Xf, idf, namesf = filter_data(X, ids, names)
    else:
        Xf, idf, namesf = X, ids, names
    st = load_mtx(path, do_filter=do_filter, device=0)
    try:
        assert st.file_shape == X.shape
        assert np.array_equal(st.kept_rows, idf)
        assert [names[j] for j in st.kept_cols] == list(namesf)
        assert st.shape == Xf.shape
        assert np.array_equal(st.to_host(), Xf)
        if name == "counts" and do_filter:
            assert 0 < st.shape[0] < X.shape[0] and 0 < st.shape[1] < X.shape[1]   # the filter did something
    finally:
        st.close()


def test_stage_cleans_nonfinite_and_mirrors_symmetric(tmp_path):
    from phet_b200 import load_mtx
    path = str(tmp_path / "n.mtx")
    with open(path, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate real symmetric\n3 3 5\n1 1 2.5\n2 1 nan\n3 1 inf\n3 2 4\n3 3 -inf\n")
    st = load_mtx(path, do_filter=False, device=0)
    try:
        want = np.array([[2.5, 0, 0], [0, 0, 4], [0, 4, 0]], dtype=np.float32)
        assert np.array_equal(st.to_host(), want)
    finally:
        st.close()


def test_fit_from_staged_matrix_equals_fit_from_numpy(tmp_path):
    """The fit that starts from HBM (StagedMatrix) returns the scores of the fit that starts from the host array."""
    from phet_b200 import PHet, load_mtx
    from oracle.ingest_oracle import filter_data, load_dataset
    rng = np.random.default_rng(5)
    N, G = 300, 500
    X = rng.poisson(1.5, (N, G)).astype(np.float64)
    X[N // 2:, :30] += rng.poisson(2.0, (N - N // 2, 30))
    path = str(tmp_path / "d_matrix.mtx")
    scipy.io.mmwrite(path, sp.coo_matrix(X))
    y = (np.arange(N) >= N // 2).astype(np.int64)
    with open(tmp_path / "d_classes.csv", "w") as fh:
        fh.write("classes\n" + "\n".join(str(int(v)) for v in y) + "\n")
    with open(tmp_path / "d_feature_names.csv", "w") as fh:
        fh.write("features\n" + "\n".join("g%d" % j for j in range(G)) + "\n")
    Xh, yh, names = load_dataset(str(tmp_path), "d")
    Xh, yh, names = filter_data(Xh, yh, names)
    np.random.seed(12345)
    a = PHet(num_subsamples=50, device=0)
    ref = a.fit_predict(Xh, yh)
    st = load_mtx(path, device=0)
    try:
        np.random.seed(12345)
        b = PHet(num_subsamples=50, device=0)
        got = b.fit_predict(st, y[st.kept_rows])
    finally:
        st.close()
    assert np.array_equal(got, ref)
    assert np.array_equal(a.ranking_, b.ranking_)


def _pandas_stable_order(s):
    import pandas as pd
    return pd.DataFrame({"score": s}).sort_values(by=["score"], ascending=False, kind="stable").index.to_numpy()


@pytest.mark.parametrize("G", [1, 2, 255, 256, 257, 2047, 2048, 2049, 4500, 30000])
@pytest.mark.parametrize("kind", ["distinct", "ties", "special"])
def test_rank_matches_pandas(G, kind):
    from phet_b200 import _capi
    rng = np.random.default_rng(G)
    if kind == "distinct":
        s = rng.permutation(G).astype(np.float64) * 1e-3 + 1e-9
    elif kind == "ties":
        s = rng.integers(0, max(2, G // 7), G).astype(np.float64) / 3.0
    else:
        s = rng.normal(size=G)
        s[rng.random(G) < 0.1] = 0.0
        s[rng.random(G) < 0.05] = -0.0
        s[rng.random(G) < 0.03] = np.inf
        s[rng.random(G) < 0.03] = -np.inf
        s[rng.random(G) < 0.05] = np.nan
    ctx = _capi.Context(0)
    try:
        order, srt = ctx.rank(s)
        want = _pandas_stable_order(s)
        assert np.array_equal(order, want)
        assert np.array_equal(srt, s[want], equal_nan=True)
        if kind == "distinct":   # the reference's default (unstable) sort agrees wherever there are no ties
            import pandas as pd
            ref = pd.DataFrame({"score": s}).sort_values(by=["score"], ascending=False).index.to_numpy()
            assert np.array_equal(order, ref)
        k = min(G, 10)
        top, _ = ctx.rank(s, k=k)
        assert np.array_equal(top, want[:k])
    finally:
        ctx.close()


def test_rank_of_golden_scores_matches_reference_ranking():
    """Reference scores of SRBCT and the reference's own sort_features output (tests/golden/make_golden_cli.py)."""
    from conftest import GOLDEN_DIR
    from phet_b200 import _capi
    from phet_b200.select import significant_features
    z = np.load(os.path.join(GOLDEN_DIR, "cli_srbct.npz"))
    scores, names = z["scores"], z["names"].tolist()
    ctx = _capi.Context(0)
    try:
        order, _ = ctx.rank(scores[:, 0])
    finally:
        ctx.close()
    assert [names[i] for i in order] == z["rank_features"].tolist()
    df = significant_features(X=scores, features_name=names, alpha=0.05, fit_type="gamma", per=95, order=order)
    assert df["features"].to_list() == z["sig_features"].tolist()
    assert np.array_equal(df["score"].to_numpy(), z["sig_scores"])


def test_cli_native_loader_and_device_ranking_equal_host_path(tmp_path):
    """main.py (C++ reader + device filter, pandas or device ranking) writes the file that the host restatement of
    train.py:77-118 (oracle/ingest_oracle.py) + the class + the pandas sort write."""
    from oracle.ingest_oracle import filter_data, load_dataset
    from phet_b200 import PHet
    from phet_b200.main import SEED, main
    from phet_b200.select import significant_features
    rng = np.random.default_rng(2)
    _write_case(tmp_path, "counts", rng)
    common = ["--file-name", "counts", "--dspath", str(tmp_path), "--methods", "phet", "--num-subsamples", "30"]
    main(common + ["--rspath", str(tmp_path / "a")])
    main(common + ["--rspath", str(tmp_path / "b"), "--ranking", "device"])
    X, y, names = load_dataset(str(tmp_path), "counts")
    X, y, names = filter_data(X, y, names)
    np.random.seed(SEED)
    scores = PHet(normalize="zscore", iqr_range=(25, 75), num_subsamples=30, feature_weight=[0.4, 0.3, 0.2, 0.1]).fit_predict(X, y)
    df = significant_features(X=scores, features_name=names, alpha=0.05, fit_type="gamma", per=95)
    os.makedirs(tmp_path / "c", exist_ok=True)
    df.to_csv(tmp_path / "c" / "counts_phet_br_features.csv", sep=",", index=False)
    a = open(tmp_path / "a" / "counts_phet_br_features.csv").read()
    b = open(tmp_path / "b" / "counts_phet_br_features.csv").read()
    c = open(tmp_path / "c" / "counts_phet_br_features.csv").read()
    assert a == c and len(a.splitlines()) > 2
    assert sorted(a.splitlines()) == sorted(b.splitlines())   # device ranking: same rows, order free inside tie groups


def test_new_entry_points_fail_loudly(tmp_path):
    """Misuse of the 8(f) entry points is an error with a message, never a silent wrong answer."""
    import ctypes as C
    from phet_b200 import _capi
    ctx = _capi.Context(0)
    try:
        with pytest.raises(_capi.PhetError, match="no scores"):
            ctx.rank()                                           # nothing resident yet
        s = np.arange(10, dtype=np.float64)
        with pytest.raises(_capi.PhetError, match="k <= G"):
            ctx.rank(s, k=11)
        with pytest.raises(_capi.PhetError, match="no staged matrix|buffer has not been produced"):
            ctx.read(_capi.BUF_STAGED, np.float32, (1,))
        y = np.repeat(np.arange(3), 20)
        X = np.random.default_rng(0).normal(size=(60, 5))
        with pytest.raises(_capi.PhetError, match="load the classes and the matrix"):
            ctx.anova(np.zeros((4, 3, 6), dtype=np.int32), 0.0)
        ctx.set_classes(np.where(y == 0)[0], np.where(y != 0)[0])
        ctx.load_matrix(X)
        bad = np.zeros((4, 3, 6), dtype=np.int32)
        bad[1, 2, 3] = 60                                        # row outside [0, N)
        with pytest.raises(_capi.PhetError, match="outside"):
            ctx.anova(bad, 0.0)
        with pytest.raises(_capi.PhetError, match="at most 64 groups"):
            ctx.kruskal(np.zeros((2, 65, 2), dtype=np.int32))
        ptr = ctx.stage_matrix(np.ascontiguousarray(X))
        with pytest.raises(_capi.PhetError, match="gene range"):
            ctx.load_matrix_genes(ptr, 5, 60, 5, _capi.PHET_F64, 3, 9)
        with pytest.raises(_capi.PhetError, match="STEP3_MIN_SIZE"):
            ctx.set_option(_capi.OPT_STEP3_MIN_SIZE, -1)
    finally:
        ctx.close()
