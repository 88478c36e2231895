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


def load_dataset(dspath: str, file_name: str):
    """train.py:47-84 for the files PHet needs.  MatrixMarket coordinate file, cells x genes,
    densified as float32 (scanpy's read_mtx default dtype)."""
    from scipy.io import mmread
    X = mmread(os.path.join(dspath, file_name + "_matrix.mtx"))
    X = np.asarray(X.toarray() if hasattr(X, "toarray") else X, dtype=np.float32)
    np.nan_to_num(X, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    y = pd.read_csv(os.path.join(dspath, file_name + "_classes.csv"), sep=",").iloc[:, 0].to_numpy().astype(np.int64)
    names = pd.read_csv(os.path.join(dspath, file_name + "_feature_names.csv"), sep=",").iloc[:, 0].to_list()
    return X, y, names


def load_dataset_device(dspath: str, file_name: str, do_filter: bool = True, device=None, minimum_samples: int = 5):
    """train.py:47-118 with the matrix never materialised on the host: C++ reader -> triplets -> B200
    (phet_b200.ingest.load_mtx); y and the feature names are filtered with the kept row / column numbers."""
    from .ingest import load_mtx
    X = load_mtx(os.path.join(dspath, file_name + "_matrix.mtx"), do_filter=do_filter, minimum_samples=minimum_samples,
                 device=device)
    y = pd.read_csv(os.path.join(dspath, file_name + "_classes.csv"), sep=",").iloc[:, 0].to_numpy().astype(np.int64)
    names = pd.read_csv(os.path.join(dspath, file_name + "_feature_names.csv"), sep=",").iloc[:, 0].to_list()
    return X, y[X.kept_rows], np.array(names)[X.kept_cols].tolist()


def filter_data(X, y, features_name, minimum_samples: int = 5):
    """train.py:88-118: cells with sum |x| >= 5; genes with CPM > 1 in >= minimum_samples cells."""
    example_sums = np.absolute(X).sum(1)
    examples_ids = np.where(example_sums >= 5)[0]
    X = X[examples_ids]
    y = y[examples_ids]
    num_examples, _ = X.shape
    temp = np.absolute(X)
    temp = (temp * 1e6) / temp.sum(axis=1).reshape((num_examples, 1))
    temp[temp > 1] = 1
    temp[temp != 1] = 0
    feature_sums = temp.sum(0)
    if num_examples <= minimum_samples or minimum_samples > num_examples // 2:
        minimum_samples = num_examples // 2
    feature_ids = np.where(feature_sums >= minimum_samples)[0]
    features_name = np.array(features_name)[feature_ids].tolist()
    return X[:, feature_ids], y, features_name


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
    p.add_argument("--permutation", type=str, default="reference", choices=["reference", "device"])
    p.add_argument("--loader", type=str, default="native", choices=["native", "scipy"],
                   help="native: multi-threaded C++ MatrixMarket reader + filter on the B200 (the matrix never "
                        "exists on the host); scipy: mmread + NumPy filter on the host (train.py:82-118 as written)")
    p.add_argument("--ranking", type=str, default="device", choices=["device", "pandas"],
                   help="device: order of the scores from phet_rank; pandas: DataFrame.sort_values (utils.py:107-136)")
    return p


def run(args) -> pd.DataFrame:
    from . import PHet
    from .select import significant_features, sort_features
    np.random.seed(seed=SEED)
    t0 = time.perf_counter()
    if args.loader == "native":
        X, y, names = load_dataset_device(args.dspath, args.file_name, not args.no_filtering, args.device)
    else:
        X, y, names = load_dataset(args.dspath, args.file_name)
        if not args.no_filtering:
            X, y, names = filter_data(X, y, names)
    print("## loader %s: %.3f s" % (args.loader, time.perf_counter() - t0))
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
