"""GPU: the reference's random draws replayed natively.

* the device Fisher-Yates (csrc/fy_apply.cu) against np.random.permutation, shared-memory and global-memory variants;
* the class default ``permutation="reference"`` (native replay inside the pipeline) against ``permutation="numpy"``
  (the same draws made by NumPy itself in Python loops) -- identical index sets, KS statistics, o, bins, scores -- and
  np.random left in the state the reference would leave it in;
* reference goldens at BASELINE's big shapes: config 2 at full width (297 x 25,475), config-4 Step-3 geometry
  (25k-cell classes: m = 158, 159 slices, last 36) and config-5 geometry (100k-cell classes: m = 316, 317 slices,
  last 144), matrices regenerated from seeds (tests/golden/big_cases.py).  Integer outputs bit-exact, scores
  rtol 1e-5 (f64) / 1e-3 (f32).
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden

sys.path.insert(0, GOLDEN_DIR)
import big_cases  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,count", [(2, 5), (3, 40), (29, 64), (252, 100), (1000, 33), (7264, 20), (25000, 9), (65536, 3),
                                     (65537, 3), (100000, 5)])
def test_device_fisher_yates_matches_numpy(n, count):
    from phet_b200 import _capi
    np.random.seed(77 + n)
    mt = _capi.MtStream.from_numpy()
    draws = mt.shuffle_draws([n], count)[0]
    ctx = _capi.Context(0)
    try:
        got = ctx.fy_apply(draws, n)
        head = ctx.fy_apply(draws, n, take=min(n, 17))
    finally:
        ctx.close()
    for r in range(count):
        ref = np.random.permutation(n)
        assert np.array_equal(got[r], ref), (n, r)
        assert np.array_equal(head[r], ref[:min(n, 17)])
    key, pos = mt.state()
    st = np.random.get_state()
    assert np.array_equal(key, st[1]) and pos == st[2]


@pytest.mark.parametrize("name", ["hbecs_like", "balanced_1200", "unbalanced_2200", "cells10k_m70", "replace_true"])
def test_native_replay_equals_numpy_draws(name):
    from phet_b200 import PHet
    g = load_golden(name)
    tag = g["dtypes"][0]
    X = np.ascontiguousarray(g["X"].astype(np.float64 if tag == "f64" else np.float32))
    y = g["y"].astype(np.int64)
    out = {}
    for mode in ("numpy", "reference"):
        np.random.seed(int(g["seed"]))
        est = PHet(**g["params"], permutation=mode, capture=True, distributed=False)
        scores = est.fit_predict(X, y)
        out[mode] = (est, scores, np.random.get_state())
    a, b = out["numpy"], out["reference"]
    assert np.array_equal(a[0].index_sets_, b[0].index_sets_)
    assert np.array_equal(a[0].H_, b[0].H_) and np.array_equal(a[0].o_, b[0].o_)
    assert np.array_equal(a[0].bins_, b[0].bins_)
    assert np.array_equal(a[1], b[1]), "same draws, same kernels: scores must be bit-identical"
    assert np.array_equal(a[2][1], b[2][1]) and a[2][2] == b[2][2], "np.random must end in the reference's state"
    assert np.array_equal(b[0].H_, g[f"H_{tag}"])


def _load_big(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    y = d["y"].astype(np.int64)
    X = big_cases.build_X(name, y)
    assert big_cases.checksum(X) == str(d["x_sha256"]), "regenerated matrix differs from the one the reference saw"
    return d, X, y


@pytest.mark.parametrize("name", list(big_cases.BIG))
@pytest.mark.parametrize("force_big", [False, True])
def test_big_geometry_matches_reference(name, force_big, monkeypatch):
    """H / o / bins / counts bit-exact, scores within tolerance, at config-2 width and config-4/5 Step-3 geometry;
    once on the default kernels and once with the global-gather Step-1 instantiation forced."""
    import json
    from phet_b200 import PHet
    if force_big:
        if name == "big_cfg5_geom":
            pytest.skip("config-5 geometry already runs the global-gather instantiation")
        monkeypatch.setenv("PHET_STEP1_FORCE_BIG", "1")
    d, X32, y = _load_big(name)
    params = json.loads(str(d["params_json"]))
    params["percentiles"] = tuple(params["percentiles"])
    for tag in big_cases.BIG[name]["dtypes"]:
        X = X32 if tag == "f32" else np.ascontiguousarray(X32.astype(np.float64))
        # capture: sorting kernels (every slice's h); else bound-and-prune, streaming (scatter) and gather formulation
        for capture, impl in ((True, "auto"), (False, "scatter"), (False, "prune")):
            monkeypatch.setenv("PHET_STEP3_IMPL", impl)
            np.random.seed(int(d["seed"]))
            est = PHet(**params, capture=capture, distributed=False)
            scores = est.fit_predict(X, y)
            if impl == "scatter":
                assert est.step3_path_ == "scatter bound-and-prune", est.step3_path_
            assert np.array_equal(est.index_sets_[:, 0], d["idx_i"]) and np.array_equal(est.index_sets_[:, 1], d["idx_j"])
            if capture:
                assert np.array_equal(est.H_, d[f"H_{tag}"]), "KS integer statistics differ"
                assert np.array_equal(est.bins_.astype(np.int64), d[f"bins_{tag}"].astype(np.int64))
                ks = d[f"P_head_{tag}"].shape[1]
                rtol = 1e-5 if tag == "f64" else 1e-3
                assert np.allclose(est.R_[:, :ks], d[f"R_head_{tag}"], rtol=rtol, atol=1e-5 if tag == "f32" else 1e-12)
                assert np.allclose(est.P_[:, :ks], d[f"P_head_{tag}"], rtol=rtol, atol=1e-300)
            assert np.array_equal(est.o_, d[f"o_{tag}"]), "o must be bit-exact"
            assert np.array_equal(est.bin_counts_, d[f"bin_counts_{tag}"])
            rtol = 1e-5 if tag == "f64" else 1e-3
            ref = d[f"scores_{tag}"]
            assert np.allclose(scores, ref, rtol=rtol, atol=1e-12), float(np.max(np.abs(scores - ref) / np.abs(ref).max()))
            if tag == "f64":
                k = min(100, scores.shape[0] // 4)
                assert np.array_equal(np.argsort(-scores[:, 0], kind="stable")[:k], np.argsort(-ref[:, 0], kind="stable")[:k])


def test_class_of_65536_cells_takes_the_32_bit_table():
    """position 65,535 is legal in a 65,536-cell class and must not meet the 16-bit table's marker."""
    from oracle import phet_oracle as po
    from phet_b200 import PHet
    rng = np.random.default_rng(5)
    n0, n1, G, S = 65536, 400, 6, 12
    y = np.zeros(n0 + n1, dtype=np.int64)
    y[n0:] = 1
    X = rng.gamma(2.0, 1.0, (n0 + n1, G)).astype(np.float32)
    np.random.seed(1)
    est = PHet(num_subsamples=S, capture=True, distributed=False)
    scores = est.fit_predict(X, y)
    idx = est.index_sets_
    np.random.seed(1)
    ref = po.fit_predict_oracle(X, y, num_subsamples=S)
    assert np.array_equal(idx[:, 0], ref.idx_i) and np.array_equal(idx[:, 1], ref.idx_j)
    assert np.array_equal(est.H_, ref.H)
    assert np.allclose(scores, ref.scores, rtol=1e-3, atol=1e-12)
