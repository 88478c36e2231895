"""Golden vectors for the multi-class branch (K >= 3).  Build container only:

    python tests/golden/make_golden_multiclass.py

Runs the UNMODIFIED reference ``PHet.fit_predict`` (oracle/ref_shim.py) and
``oracle.multiclass_oracle.fit_predict_multiclass_oracle`` on the same seed, REQUIRES
max|oracle - reference| == 0.0, and stores inputs, reference scores and the oracle's per-stage intermediates in
``tests/golden/multi_<case>.npz``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import multiclass_oracle as mo  # noqa: E402
from oracle.ref_shim import load_reference_phet  # noqa: E402

SEED = 12345


def synth(sizes, G, seed, kind="gamma"):
    rng = np.random.default_rng(seed)
    y = np.concatenate([np.full(n, c) for c, n in enumerate(sizes)]).astype(np.int64)
    N = len(y)
    if kind == "gamma":
        X = rng.gamma(2.0, 1.0, (N, G))
    else:
        X = (rng.poisson(1.5, (N, G)) * (rng.random((N, G)) > 0.6)).astype(np.float64)
    k = max(G // 10, 1)
    for c in range(len(sizes)):
        X[y == c, :k] += 0.6 * c                                   # mean shifts across classes
    mu = X[:, k:2 * k].mean(0)
    X[y == 1, k:2 * k] = (X[y == 1, k:2 * k] - mu) * 2.5 + mu      # one class more dispersed
    X[:, 2 * k] = 3.0                                              # constant gene
    X[y == 0, 2 * k + 1] = 1.0                                     # constant inside each class, different between
    X[y == 1, 2 * k + 1] = 2.0
    X[y >= 2, 2 * k + 1] = 4.0
    shuffle = rng.permutation(N)                                   # classes interleaved
    return X[shuffle].astype(np.float32), y[shuffle]


CASES = [
    # name, sizes, G, S, dtype, kwargs
    ("k3_small", (40, 55, 35), 90, 40, "f64", {}),                                  # m = 35 (sqrt < 8 -> min size)
    ("k3_sqrt", (150, 100, 120), 60, 30, "f64", {}),                                # m = 10, 10 slices
    ("k4_f32", (90, 70, 100, 64), 50, 25, "f32", {}),                               # m = 8, float32 arithmetic
    ("k3_pairwise_mean", (70, 64, 80), 40, 20, "f64", dict(multiconditions_strategy="pairwise",
                                                             multiconditions_aggregation="mean")),
    ("k3_counts_p1090", (81, 90, 100), 60, 30, "f64", dict(percentiles=(10, 90), kind="counts")),
    ("k3_unbalanced", (2500, 160, 1400), 24, 10, "f64", {}),      # m = 12; last slices of 1404 / 4 / 1244 cells
    ("k3_oneslice_big", (700, 40, 650), 20, 10, "f32", {}),       # m = 40 = min_size: one slice of 690 / 40 / 650 cells
    ("k3_kruskal", (150, 100, 120), 60, 30, "f64", dict(group_test="kruskal")),              # m = 10 > 5 -> kruskal
    ("k4_kruskal_counts_f32", (90, 70, 100, 64), 50, 25, "f32", dict(group_test="kruskal", kind="counts")),  # heavy ties
    ("k3_kruskal_m37", (40, 55, 37), 40, 20, "f64", dict(group_test="kruskal")),            # n = 111: partial last pass
]


def main():
    PHet = load_reference_phet()
    for name, sizes, G, S, dt, kw in CASES:
        kw = dict(kw)
        X, y = synth(sizes, G, seed=len(name), kind=kw.pop("kind", "gamma"))
        X = X.astype(np.float64 if dt == "f64" else np.float32)
        percentiles = kw.pop("percentiles", (25, 75))
        strategy = kw.get("multiconditions_strategy", "one_vs_all")
        agg = kw.get("multiconditions_aggregation", "max")
        gt = kw.get("group_test", "anova")
        np.random.seed(SEED)
        ref = PHet(normalize="zscore", percentiles=percentiles, num_subsamples=S, subsampling_size="sqrt",
                   delta_type="iqr", adjust_pvalue=False, feature_weight=[0.4, 0.3, 0.2, 0.1],
                   **dict(dict(group_test="anova"), **kw)).fit_predict(X=np.array(X, copy=True), y=y)
        np.random.seed(SEED)
        o = mo.fit_predict_multiclass_oracle(X, y, percentiles=percentiles, num_subsamples=S, strategy=strategy,
                                             aggregation=agg, group_test=gt)
        diff = float(np.max(np.abs(o["scores"] - ref)))
        print("%-20s m=%d K=%d combos=%d slices=%d  max|oracle-reference| = %g" % (name, o["m"], o["K"], len(o["combos"]),
                                                                                 len(o["slices"]), diff))
        assert diff == 0.0, name
        np.random.seed(SEED)
        o2 = mo.fit_predict_multiclass_oracle(X, y, percentiles=percentiles, num_subsamples=S, strategy=strategy,
                                              aggregation=agg, library=False, group_test=gt)
        print("    plain-formula flavour: max rel diff P = %g, scores = %g" % (
            float(np.max(np.abs(o2["P"] - o["P"]) / np.maximum(o["P"], 1e-300))), float(np.max(np.abs(o2["scores"] - ref)))))
        C = len(o["combos"])
        np.savez_compressed(
            os.path.join(HERE, "multi_%s.npz" % name), X=X, y=y.astype(np.int8), S=np.int64(S), m=np.int64(o["m"]),
            percentiles=np.array(percentiles), strategy=np.array(strategy), aggregation=np.array(agg),
            group_test=np.array(gt),
            scores_ref=ref, idx_anova=o["idx_anova"].astype(np.int32), P=o["P"], fsum=o["fsum"], rsum=o["rsum"],
            r=o["r"], f=o["f"], H=o["H"], o_combo=o["o_combo"], o=o["o"], bins=o["bins"].astype(np.int8),
            edges=o["edges"], bin_counts=o["bin_counts"], min_size=np.int64(o["min_size"]),
            **{"idx_i_%d" % c: o["idx_combo"][c][0].astype(np.int32) for c in range(C)},
            **{"idx_j_%d" % c: o["idx_combo"][c][1].astype(np.int32) for c in range(C)},
            n_last=np.array([sl[-1][1] for sl in o["slices_c"]], dtype=np.int64),
            **{"perm_i_%d" % c: o["perms"][c][0][:, :sum(n for _, n in o["slices_c"][c])].astype(np.uint32) for c in range(C)},
            **{"perm_j_%d" % c: o["perms"][c][1][:, :sum(n for _, n in o["slices_c"][c])].astype(np.uint32) for c in range(C)})


if __name__ == "__main__":
    main()
