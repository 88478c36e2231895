"""Generate the committed golden vectors by running the UNMODIFIED reference PHet.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it (1) seeds the legacy global RNG exactly as ``src/train.py:26`` does,
(2) runs ``/root/reference/src/model/phet.py`` ``PHet(...).fit_predict(X, y)`` under
``oracle/ref_shim.py``, (3) re-seeds and runs ``oracle.phet_oracle.fit_predict_oracle`` and
REQUIRES max|scores_oracle - scores_reference| == 0.0, then (4) stores inputs, the reference
scores and the oracle's per-stage intermediates (which the reference does not expose) in
``tests/golden/<case>.npz``.  Library versions are recorded in ``tests/golden/VERSIONS.json``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import phet_oracle as po  # noqa: E402
from oracle.ref_shim import load_reference_phet  # noqa: E402

SEED = 12345  # src/train.py:26
KEEP_S = 8    # subsamples for which P[:, s] and delta[:, s] are stored


def load_srbct():
    from scipy.io import mmread
    X = np.asarray(mmread("/root/reference/samples/srbct_matrix.mtx").toarray(), dtype=np.float64)
    y = np.loadtxt("/root/reference/samples/srbct_classes.csv", skiprows=1).astype(np.int64)
    return X, y


def hbecs_classes():
    return np.loadtxt("/root/reference/samples/hbecs_classes.csv", skiprows=1).astype(np.int64)


def synth(N, G, n1, seed, kind="gamma"):
    """Synthetic expression matrix, float32-exact values.  First rows control, then case."""
    rng = np.random.default_rng(seed)
    y = np.zeros(N, dtype=np.int64)
    y[N - n1:] = 1
    if kind == "gamma":
        X = rng.gamma(2.0, 1.0, (N, G))
    elif kind == "counts":  # zero-inflated counts: heavy ties, like scRNA-seq
        X = rng.poisson(1.5, (N, G)) * (rng.random((N, G)) > 0.7)
        X = X.astype(np.float64)
    else:
        raise ValueError(kind)
    k = max(G // 20, 1)
    X[y == 1, :k] += 1.0                                   # mean-shifted genes
    mu = X[:, k:2 * k].mean(0)
    X[y == 1, k:2 * k] = (X[y == 1, k:2 * k] - mu) * 2 + mu  # dispersion-shifted genes
    return X.astype(np.float32), y


CASES = []


def case(name, X, y, dtypes=("f64", "f32"), **kw):
    CASES.append((name, X, y, dtypes, kw))


def build_cases():
    Xs, ys = load_srbct()
    case("srbct_full", Xs, ys, dtypes=("f64",), num_subsamples=1000)          # BASELINE config 1
    case("srbct_s200", Xs[:, :600], ys, num_subsamples=200)
    yh = hbecs_classes()                                                       # 252 / 45 -> m=45, z
    Xh, _ = synth(len(yh), 300, int(yh.sum()), 1, kind="counts")
    # rows must follow the real class vector, not "controls first"
    order = np.argsort(yh, kind="stable")
    Xh2 = np.empty_like(Xh)
    Xh2[order] = Xh
    Xh2[:, 7] = 3.0                                                            # constant gene -> NaN path
    Xh2[:, 11] = yh * 100.0 + (np.arange(len(yh)) % 3) * 0.25                  # |z| > 37.68 -> p == 0 dropped
    case("hbecs_like", Xh2, yh, num_subsamples=100)
    Xb, yb = synth(1200, 200, 600, 2)                                          # m=24 (t), 25 slices
    case("balanced_1200", Xb, yb, num_subsamples=60)
    Xu, yu = synth(2200, 120, 1000, 3)                                         # 1200/1000 -> m=31 (z), last n=8
    case("unbalanced_2200", Xu, yu, num_subsamples=60)
    Xl, yl = synth(10000, 48, 5000, 4)                                         # m=70: lerp 17.25/51.75, 72 slices
    case("cells10k_m70", Xl, yl, num_subsamples=40)
    case("cells10k_p1090_b3", Xl[:, :32], yl, num_subsamples=30, percentiles=(10, 90),
         feature_weight=[0.5, 0.3, 0.2])
    Xr, yr = synth(40, 50, 15, 5)                                              # int size > class -> replace=True
    case("replace_true", Xr, yr, num_subsamples=50, subsampling_size=20)
    Xo, yo = synth(600, 80, 250, 6)                                            # "optimum" size
    case("optimum_600", Xo, yo, num_subsamples=40, subsampling_size="optimum")
    Xn, yn = synth(400, 64, 200, 7, kind="counts")                             # normalize=None
    case("nonorm_400", Xn, yn, num_subsamples=40, normalize=None)
    Xw, yw = synth(30000, 8, 15000, 8)                                         # m=122, E=4 sort width
    case("cells30k_m122", Xw, yw, dtypes=("f64",), num_subsamples=16)


def run_case(PHet, name, X, y, dtypes, kw):
    out = {"y": y.astype(np.int8), "seed": np.int64(SEED)}
    store_dtype = np.float32 if np.array_equal(X.astype(np.float32).astype(np.float64),
                                               X.astype(np.float64)) else np.float64
    out["X"] = X.astype(store_dtype)
    params = dict(normalize="zscore", percentiles=(25, 75), num_subsamples=1000,
                  subsampling_size="sqrt", feature_weight=None)
    params.update(kw)
    out["params_json"] = np.array(json.dumps(params))
    for tag in dtypes:
        dt = np.float64 if tag == "f64" else np.float32
        Xd = np.ascontiguousarray(X.astype(dt))
        np.random.seed(SEED)
        ref = PHet(**params).fit_predict(np.array(Xd, copy=True), y)
        np.random.seed(SEED)
        res = po.fit_predict_oracle(np.array(Xd, copy=True), y, **params)
        diff = float(np.max(np.abs(ref - res.scores)))
        assert diff == 0.0, (name, tag, diff)
        assert ref.dtype == np.float64 and ref.shape == (X.shape[1], 1)
        out[f"scores_{tag}"] = ref
        for k in ("fsum", "rsum", "f", "r", "o", "edges"):
            out[f"{k}_{tag}"] = getattr(res, k)
        out[f"H_{tag}"] = res.H.astype(np.int32)
        out[f"bins_{tag}"] = res.bins.astype(np.int8)
        out[f"bin_counts_{tag}"] = res.bin_counts
        ks = min(KEEP_S, params["num_subsamples"])
        out[f"P_head_{tag}"] = res.P[:, :ks]
        out[f"R_head_{tag}"] = res.R[:, :ks]
        if "idx_i" not in out:  # draws do not depend on dtype
            out["m"] = np.int64(res.m)
            out["replace"] = np.bool_(res.replace)
            out["idx_i"] = res.idx_i.astype(np.int32)
            out["idx_j"] = res.idx_j.astype(np.int32)
            out["perm_i"] = res.extra["perm_i"].astype(np.uint16)
            out["perm_j"] = res.extra["perm_j"].astype(np.uint16)
            out["slices"] = np.array(res.slices, dtype=np.int64)
        else:
            assert np.array_equal(out["idx_i"], res.idx_i) and np.array_equal(out["perm_j"], res.extra["perm_j"])
        print(f"  {name:22s} {tag}: G={X.shape[1]} N={X.shape[0]} m={res.m} replace={res.replace} "
              f"slices={len(res.slices)} bins={res.bin_counts.tolist()} sum={ref.sum():.12f} ref==oracle")
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def main():
    import scipy
    import sklearn
    PHet = load_reference_phet()
    build_cases()
    only = set(sys.argv[1:])
    for name, X, y, dtypes, kw in CASES:
        if only and name not in only:
            continue
        run_case(PHet, name, X, y, dtypes, kw)
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as fh:
        json.dump({"numpy": np.__version__, "scipy": scipy.__version__, "sklearn": sklearn.__version__,
                   "python": sys.version.split()[0], "seed": SEED,
                   "reference": "/root/reference/src/model/phet.py (unmodified, via oracle/ref_shim.py)"}, fh, indent=1)


if __name__ == "__main__":
    main()
