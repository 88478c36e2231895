"""Inputs of the three "big geometry" golden cases.  The matrices are too large
to commit (30 / 13 / 13 MB), so the fixtures store the seeds and a checksum; the generator
(``make_golden_big.py``, build container, runs the unmodified reference) and the tests (GPU box) both rebuild
X from here.  ``numpy.random.default_rng`` streams are stable for a numpy version; VERSIONS.json pins it and
the checksum catches a drift.

    big_hbecs_fullwidth   297 x 25,475 (BASELINE config 2's shape; the real matrix is absent from the reference
                          checkout), zero-inflated counts, the REAL hbecs_classes.csv class
                          vector (252 / 45 -> m = 45, z branch, 1 slice), S = 1000
    big_cfg4_geom         50,000 x 64, 25k / 25k -> m = 158, 159 slices, last 36 (BASELINE config 4's Step-3
                          geometry), S = 1000, float32
    big_cfg5_geom         200,000 x 16, 100k / 100k -> m = 316, 317 slices, last 144 (config 5's geometry), S = 50
"""
from __future__ import annotations

import hashlib

import numpy as np

SEED = 12345  # src/train.py:26

# class vector of samples/hbecs_classes.csv (252 zeros, 45 ones), run-length encoded so that the test does not
# need the reference checkout: filled in by make_golden_big.py into the fixture itself ("y").

BIG = {
    "big_hbecs_fullwidth": dict(N=297, G=25475, kind="counts", xseed=101, S=1000, dtypes=("f32", "f64")),
    "big_cfg4_geom": dict(N=50000, G=64, kind="gamma", xseed=102, S=1000, dtypes=("f32",)),
    "big_cfg5_geom": dict(N=200000, G=16, kind="gamma", xseed=103, S=50, dtypes=("f32",)),
}


def build_X(name: str, y: np.ndarray) -> np.ndarray:
    """float32 matrix (cells x genes) of a big case, rows following the class vector ``y``."""
    spec = BIG[name]
    N, G = spec["N"], spec["G"]
    assert len(y) == N
    rng = np.random.default_rng(spec["xseed"])
    if spec["kind"] == "gamma":
        X = rng.gamma(2.0, 1.0, (N, G)).astype(np.float32)
    else:  # zero-inflated counts: heavy ties, like scRNA-seq
        X = (rng.poisson(1.5, (N, G)) * (rng.random((N, G)) > 0.7)).astype(np.float32)
    case = y == 1
    k = max(G // 20, 1)
    X[case, :k] += 1.0                                           # mean-shifted genes
    mu = X[:, k:2 * k].mean(0, dtype=np.float64).astype(np.float32)
    X[case, k:2 * k] = (X[case, k:2 * k] - mu) * 2 + mu          # dispersion-shifted genes
    if spec["kind"] == "counts":
        X[:, 7] = 3.0                                            # constant gene -> NaN path
        X[:, 11] = y * 100.0 + (np.arange(N) % 3) * 0.25         # |z| > 37.68 -> p == 0, dropped from Fisher's sum
    return np.ascontiguousarray(X)


def checksum(X: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest()


def default_y(name: str) -> np.ndarray:
    spec = BIG[name]
    y = np.zeros(spec["N"], dtype=np.int64)
    y[spec["N"] // 2:] = 1
    return y
