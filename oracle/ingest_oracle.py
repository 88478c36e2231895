"""Host restatement of the reference's input path, ``/root/reference/src/train.py:77-118`` (test infrastructure).

TEST INFRASTRUCTURE ONLY: the checker of the library's MatrixMarket reader + device filter (csrc/mtx.cu).  Nothing
under ``phet_b200/`` imports this module.

* ``load_dataset``: train.py:77-85 -- ``sc.read_mtx`` is ``scipy.io.mmread(...)`` densified as float32 (anndata / scanpy
  are absent from this image, so this leg is pinned to scipy's reader, not to scanpy's wrapper), ``np.nan_to_num``,
  the ``classes`` / ``features`` columns of the two CSV files.
* ``filter_data``: train.py:88-118 line for line -- cells with sum |x| >= 5; genes whose counts per million exceed 1
  in at least ``minimum_samples`` kept cells (lowered to cells // 2 under the reference's rule).
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd


def load_dataset(dspath: str, file_name: str):
    from scipy.io import mmread
    X = mmread(os.path.join(dspath, file_name + "_matrix.mtx"))
    X = np.asarray(X.toarray() if hasattr(X, "toarray") else X, dtype=np.float32)
    np.nan_to_num(X, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    y = pd.read_csv(os.path.join(dspath, file_name + "_classes.csv"), sep=",")["classes"].to_numpy().astype(np.int64)
    names = pd.read_csv(os.path.join(dspath, file_name + "_feature_names.csv"), sep=",")["features"].to_list()
    return X, y, names


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
