"""Golden vectors of the sibling IQR scorers, from the UNMODIFIED reference classes.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_siblings.py

Runs ``/root/reference/src/model/hvf.py`` ``HIQR`` (both ``per_condition`` settings) and
``/root/reference/src/model/deltaiqrmean.py`` ``DeltaIQRMean`` (both ``calculate_deltamean`` settings),
requires ``oracle.siblings_oracle`` to reproduce them with max abs difference 0.0, and stores inputs +
reference outputs + the oracle's per-class intermediates in ``tests/golden/siblings_<case>.npz``.
(``hvf.py`` imports anndata / scanpy at module top for its other class; empty stubs as in oracle/ref_shim.py.)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import siblings_oracle as so  # noqa: E402

for name in ("anndata", "scanpy"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference/src")
from model.deltaiqrmean import DeltaIQRMean  # noqa: E402
from model.hvf import HIQR  # noqa: E402


def cases():
    rng = np.random.default_rng(42)
    out = {}
    X = rng.gamma(2.0, 1.0, (300, 40))
    y = np.r_[np.zeros(170, dtype=np.int64), np.ones(130, dtype=np.int64)]
    X[y == 1, :4] += 1.0
    X[y == 1, 4:8] *= 2.0
    out["gamma_f64_none"] = (X, y, None, (25, 75))
    out["gamma_f64_zscore"] = (X, y, "zscore", (25, 75))
    Xc = (rng.poisson(1.5, (500, 36)) * (rng.random((500, 36)) > 0.6)).astype(np.float32)
    Xc[:, 5] = 3.0                                   # constant gene: NaN under zscore in the reference
    yc = rng.permutation(np.r_[np.zeros(260, dtype=np.int64), np.ones(240, dtype=np.int64)])  # interleaved classes
    out["counts_f32_zscore"] = (Xc, yc, "zscore", (25, 75))
    out["counts_f32_none"] = (Xc, yc, None, (25, 75))
    Xu = rng.normal(0, 1, (2200, 30))
    yu = np.r_[np.zeros(1200, dtype=np.int64), np.ones(1000, dtype=np.int64)]
    Xu[yu == 1, :5] = Xu[yu == 1, :5] * 3 - 0.5
    out["normal_p1090_zscore"] = (Xu, yu, "zscore", (10, 90))
    out["normal_f32_p0595_none"] = (Xu.astype(np.float32), yu, None, (5, 95))
    return out


def main():
    for name, (X, y, normalize, rng_) in cases().items():
        rec = dict(X=X, y=y, normalize=np.array("none" if normalize is None else normalize), iqr_range=np.array(rng_))
        for pc in (False, True):
            ref = HIQR(per_condition=pc, normalize=normalize, iqr_range=rng_).fit_predict(X.copy(), y)
            got = so.hiqr(X.copy(), y, per_condition=pc, normalize=normalize, iqr_range=rng_)
            assert ref.shape == got.shape and np.max(np.abs(ref - got)) == 0.0, (name, "HIQR", pc)
            rec["hiqr_pc%d" % pc] = ref
        for dm in (False, True):
            ref = DeltaIQRMean(calculate_deltamean=dm, normalize=normalize, iqr_range=rng_).fit_predict(X.copy(), y)
            got = so.delta_iqr_mean(X.copy(), y, calculate_deltamean=dm, normalize=normalize, iqr_range=rng_)
            assert ref.shape == got.shape and np.max(np.abs(ref - got)) == 0.0, (name, "DeltaIQRMean", dm, np.max(np.abs(ref - got)))
            rec["dim_dm%d" % dm] = ref
        st = so.class_stats(X.copy(), y, normalize, rng_)
        for k in ("iqr0", "iqr1", "iqr_all", "mean0", "mean1", "ss0", "ss1"):
            rec["st_" + k] = np.asarray(st[k], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, "siblings_%s.npz" % name), **rec)
        print(name, "ok", X.shape, X.dtype)


if __name__ == "__main__":
    main()
