"""Golden vectors at BASELINE's big shapes, from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_big.py [case ...]

Same protocol as make_golden.py (seed 12345 -> reference PHet.fit_predict under oracle/ref_shim.py; re-seed ->
oracle; REQUIRE max|oracle - reference| == 0.0; store reference scores + the oracle's per-stage
intermediates), except that the matrix is NOT stored: ``big_cases.build_X`` regenerates it from a seed and the
fixture keeps its SHA-256.  The permutations are not stored either (they are 2 x G x n positions); the tests
replay them from ``np.random`` -- which is exactly what the product's reference mode has to get right.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import big_cases  # noqa: E402
from oracle import phet_oracle as po  # noqa: E402
from oracle.ref_shim import load_reference_phet  # noqa: E402

KEEP_S = 8


def run(PHet, name):
    spec = big_cases.BIG[name]
    if name == "big_hbecs_fullwidth":
        y = np.loadtxt("/root/reference/samples/hbecs_classes.csv", skiprows=1).astype(np.int64)
    else:
        y = big_cases.default_y(name)
    X32 = big_cases.build_X(name, y)
    params = dict(normalize="zscore", percentiles=(25, 75), num_subsamples=spec["S"], subsampling_size="sqrt",
                  feature_weight=None)
    out = {"y": y.astype(np.int8), "seed": np.int64(big_cases.SEED), "x_sha256": np.array(big_cases.checksum(X32)),
           "params_json": np.array(json.dumps(params))}
    for tag in spec["dtypes"]:
        dt = np.float32 if tag == "f32" else np.float64
        Xd = np.ascontiguousarray(X32.astype(dt))
        t0 = time.perf_counter()
        np.random.seed(big_cases.SEED)
        ref = PHet(**params).fit_predict(np.array(Xd, copy=True), y)
        t_ref = time.perf_counter() - t0
        np.random.seed(big_cases.SEED)
        res = po.fit_predict_oracle(np.array(Xd, copy=True), y, **params)
        diff = float(np.max(np.abs(ref - res.scores)))
        assert diff == 0.0, (name, tag, diff)
        out[f"scores_{tag}"] = ref
        out[f"ref_seconds_{tag}"] = np.float64(t_ref)
        for k in ("fsum", "rsum", "f", "r", "o", "edges"):
            out[f"{k}_{tag}"] = getattr(res, k)
        out[f"H_{tag}"] = res.H.astype(np.int32)
        out[f"bins_{tag}"] = res.bins.astype(np.int8)
        out[f"bin_counts_{tag}"] = res.bin_counts
        out[f"P_head_{tag}"] = res.P[:, :KEEP_S]
        out[f"R_head_{tag}"] = res.R[:, :KEEP_S]
        if "idx_i" not in out:
            out["m"] = np.int64(res.m)
            out["idx_i"] = res.idx_i.astype(np.int32)
            out["idx_j"] = res.idx_j.astype(np.int32)
            out["slices"] = np.array(res.slices, dtype=np.int64)
            # spot values of the permutation stream, so a replay can be checked without storing it all
            pi, pj = res.extra["perm_i"], res.extra["perm_j"]
            out["perm_i_first"] = pi[0].astype(np.uint32)
            out["perm_j_last"] = pj[-1].astype(np.uint32)
            out["perm_i_colsum"] = pi.astype(np.int64).sum(0) % (1 << 31)
        print(f"  {name:22s} {tag}: N={X32.shape[0]} G={X32.shape[1]} m={res.m} slices={len(res.slices)} "
              f"bins={res.bin_counts.tolist()} sum={ref.sum():.12f} ref {t_ref:.1f} s  ref==oracle", flush=True)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def main():
    PHet = load_reference_phet()
    names = sys.argv[1:] or list(big_cases.BIG)
    for n in names:
        run(PHet, n)


if __name__ == "__main__":
    main()
