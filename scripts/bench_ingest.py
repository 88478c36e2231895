"""Measurement of the two 8(f) rows either side of the fit (run on the GPU box):
input path (MatrixMarket -> filtered matrix in HBM) against scipy mmread + the NumPy filter of train.py:88-118,
and phet_rank against pandas' descending sort.  Prints one JSON line per measurement."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phet_b200 import _capi, load_mtx  # noqa: E402
from oracle.ingest_oracle import filter_data  # noqa: E402


def write_mtx(path, N, G, density, rng):
    nnz = int(N * G * density)
    flat = np.unique(rng.integers(0, N * G, nnz))
    r, c = flat // G + 1, flat % G + 1
    v = rng.poisson(2.0, flat.size) + 1
    with open(path, "w") as fh:
        fh.write("%%%%MatrixMarket matrix coordinate integer general\n%d %d %d\n" % (N, G, flat.size))
    pd.DataFrame({"r": r, "c": c, "v": v}).to_csv(path, sep=" ", header=False, index=False, mode="a")
    return flat.size


def main():
    N, G, density = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (5000, 20000, 0.1)
    rng = np.random.default_rng(0)
    d = tempfile.mkdtemp()
    path = os.path.join(d, "big_matrix.mtx")
    nnz = write_mtx(path, N, G, density, rng)
    size = os.path.getsize(path)
    # host path (train.py:82-118 as main.py restates it)
    from scipy.io import mmread
    t0 = time.perf_counter()
    X = mmread(path)
    X = np.asarray(X.toarray() if hasattr(X, "toarray") else X, dtype=np.float32)
    np.nan_to_num(X, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
    t1 = time.perf_counter()
    Xf, idf, _ = filter_data(X, np.arange(N), ["g%d" % j for j in range(G)])
    t2 = time.perf_counter()
    # native path
    best = None
    for it in range(3):
        t3 = time.perf_counter()
        st = load_mtx(path, device=0)
        t4 = time.perf_counter()
        if best is None or t4 - t3 < best[0]:
            best = (t4 - t3, dict(st.timings))
        same = bool(np.array_equal(st.kept_rows, idf) and st.shape == Xf.shape and (it > 0 or np.array_equal(st.to_host(), Xf)))
        st.close()
    print(json.dumps({"what": "input path: MatrixMarket -> filtered float32 matrix", "cells": N, "genes": G, "entries": nnz,
                      "file_mb": size / 1e6, "host_scipy_mmread_dense_s": t1 - t0, "host_numpy_filter_s": t2 - t1,
                      "native_total_s": best[0], "native_parse_s": best[1]["parse_s"], "native_stage_s": best[1]["stage_s"],
                      "parse_mb_per_s": size / 1e6 / best[1]["parse_s"], "identical": same,
                      "host_cores": os.cpu_count(), "kept": list(Xf.shape)}))
    ctx = _capi.Context(0)
    for Gr in (30000, 60000, 200000):
        s = rng.gamma(2.0, 1.0, Gr)
        ctx.rank(s)
        t0 = time.perf_counter()
        for _ in range(5):
            order, _ = ctx.rank(s)
        t1 = time.perf_counter()
        df = pd.DataFrame({"score": s})
        t2 = time.perf_counter()
        for _ in range(5):
            ref = df.sort_values(by=["score"], ascending=False).index.to_numpy()
        t3 = time.perf_counter()
        print(json.dumps({"what": "phet_rank (H2D + 2 kernels + D2H) vs pandas sort_values", "G": Gr,
                          "device_ms": (t1 - t0) / 5 * 1e3, "pandas_ms": (t3 - t2) / 5 * 1e3,
                          "identical": bool(np.array_equal(order, ref))}))
    ctx.close()


if __name__ == "__main__":
    main()
