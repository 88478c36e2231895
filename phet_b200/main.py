"""CLI entry point for the PHet method, mirroring ``/root/reference/src/main.py`` +
``src/train.py`` for ``--methods phet_br`` only (``phet`` is accepted as an alias).

    python -m phet_b200.main --file-name srbct --dspath samples --rspath out --methods phet \\
        --num-subsamples 1000 --iqr-range 25 75 --feature-weight 0.4 0.3 0.2 0.1

Same flags and defaults as the reference for this path (main.py:119-232).  What it does, with the
reference line it follows: seed NumPy with 12345 (train.py:26); load ``<file-name>_matrix.mtx`` as
float32 cells x genes (train.py:82-84), ``_classes.csv`` and ``_feature_names.csv``; unless
``--no-filtering``, drop cells with |x| row-sum < 5 and genes with CPM > 1 in fewer than 5 cells
(train.py:88-118); run ``PHet.fit_predict`` on the B200 (train.py:474-477); select and rank
(train.py:496-504); write ``<file-name>_phet_br_features.csv``.  The reference's other methods, UMAP,
clustering and plots are out of scope.  The three defects that stop the shipped ``main.py`` from
running at all are shimmed, not reproduced.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import pandas as pd

SEED = 12345  # train.py:26


def _read_labels(dspath: str, file_name: str):
    """train.py:77-81: the ``classes`` column of <file-name>_classes.csv and the ``features`` column of
    <file-name>_feature_names.csv (named columns, as the reference reads them)."""
    yf = pd.read_csv(os.path.join(dspath, file_name + "_classes.csv"), sep=",")
    nf = pd.read_csv(os.path.join(dspath, file_name + "_feature_names.csv"), sep=",")
    if "classes" not in yf.columns or "features" not in nf.columns:
        raise KeyError("expected a 'classes' column in %s_classes.csv and a 'features' column in %s_feature_names.csv "
                       "(train.py:79,81)" % (file_name, file_name))
    return yf["classes"].to_numpy().astype(np.int64), nf["features"].to_list()


def load_dataset_device(dspath: str, file_name: str, do_filter: bool = True, device=None, minimum_samples: int = 5):
    """train.py:47-118 with the matrix never materialised on the host: C++ reader -> triplets -> B200
    (phet_b200.ingest.load_mtx); y and the feature names are filtered with the kept row / column numbers."""
    from .ingest import load_mtx
    X = load_mtx(os.path.join(dspath, file_name + "_matrix.mtx"), do_filter=do_filter, minimum_samples=minimum_samples,
                 device=device)
    y, names = _read_labels(dspath, file_name)
    return X, y[X.kept_rows], np.array(names)[X.kept_cols].tolist()


def build_parser():
    p = argparse.ArgumentParser(description="PHet on B200 (drop-in for main.py --methods phet_br).")
    p.add_argument("--dspath", type=str, default=os.getcwd())
    p.add_argument("--rspath", type=str, default=os.getcwd())
    p.add_argument("--file-name", type=str, required=True)
    p.add_argument("--methods", nargs="+", type=str, default=["phet_br"], choices=["phet", "phet_br"])
    p.add_argument("--num-jobs", type=int, default=2, help="accepted for compatibility; PHet never used it")
    p.add_argument("--no-filtering", action="store_true", default=False)
    p.add_argument("--iqr-range", nargs="+", type=float, default=(25, 75))
    p.add_argument("--normalize", type=str, default="zscore")
    p.add_argument("--num-subsamples", type=int, default=1000)
    p.add_argument("--feature-weight", nargs="+", type=float, default=[0.4, 0.3, 0.2, 0.1])
    p.add_argument("--alpha", type=float, default=0.05)
    p.add_argument("--per", type=int, default=95)
    p.add_argument("--fit-type", type=str, default="gamma", choices=["expon", "gamma", "lognormal", "percentile"])
    p.add_argument("--no-sort-by-pvalue", action="store_true", default=False)
    p.add_argument("--top-k-features", type=int, default=100)
    p.add_argument("--device", type=int, default=None)
    p.add_argument("--permutation", type=str, default="reference", choices=["reference", "numpy", "device"])
    p.add_argument("--ranking", type=str, default="pandas", choices=["device", "pandas"],
                   help="pandas (default): DataFrame.sort_values exactly as utils.py:107-136 (its quicksort decides the "
                        "order inside a group of EQUAL scores, as in the reference); device: order from phet_rank "
                        "(equal scores keep ascending gene order)")
    return p


def run(args) -> pd.DataFrame:
    from . import PHet
    from .select import significant_features, sort_features
    np.random.seed(seed=SEED)
    if args.normalize != "zscore":
        # the reference's PHet call hard-codes normalize="zscore" (train.py:474); --normalize only reaches HIQR / DeltaIQRMean
        print("## --normalize %s is not forwarded to PHet (train.py:474 passes normalize='zscore'); ignored as in the reference" % args.normalize)
    if args.top_k_features != 100:
        print("## --top-k-features only feeds the reference's plots and benchmarks (train.py:136-157, 509-591): out of scope here")
    t0 = time.perf_counter()
    # MatrixMarket file -> filtered float32 matrix in HBM (C++ reader + device filter; replaces train.py:82-118)
    X, y, names = load_dataset_device(args.dspath, args.file_name, not args.no_filtering, args.device)
    print("## loader: %.3f s" % (time.perf_counter() - t0))
    print("## %s: %d examples x %d features; classes %s" % (args.file_name, X.shape[0], X.shape[1],
                                                            np.bincount(y).tolist()))
    t0 = time.perf_counter()
    estimator = PHet(normalize="zscore", iqr_range=tuple(args.iqr_range), num_subsamples=args.num_subsamples,
                     subsampling_size="sqrt", delta_type="iqr", adjust_pvalue=False,
                     feature_weight=args.feature_weight, device=args.device, permutation=args.permutation)
    scores = estimator.fit_predict(X=X, y=y)
    print("## PHet fit_predict: %.3f s (host draws %.3f s), %d kernel launches" % (
        time.perf_counter() - t0, estimator.timings_.get("host_draws_s", 0.0), estimator.launches_))
    order = estimator.ranking_ if args.ranking == "device" else None
    if not args.no_sort_by_pvalue:
        df = significant_features(X=scores, features_name=names, alpha=args.alpha, fit_type=args.fit_type,
                                  per=args.per, order=order)
    else:
        df = sort_features(X=scores, features_name=names, order=order)
    os.makedirs(args.rspath, exist_ok=True)
    out = os.path.join(args.rspath, args.file_name + "_phet_br_features.csv")
    df.to_csv(out, sep=",", index=False)
    print("## %d features written to %s; top: %s" % (len(df), out, df["features"].head(5).tolist()))
    return df


def main(argv=None):
    args = build_parser().parse_args(argv)
    run(args)


if __name__ == "__main__":
    main(sys.argv[1:])
