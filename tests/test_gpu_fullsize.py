"""GPU: BASELINE.json's full-size configurations (configs[2]: 10k x 30k, configs[3]: 50k x 30k; S = 1000) checked
through size-independent properties -- the oracle cannot run a whole fit at these sizes (hours on a CPU):

* the scores are finite, non-negative and sum to 2 (phet.py:510-516: two normalised terms), bins partition the genes;
* a second run is bit-identical (fixed-order reductions everywhere);
* genes are independent: a fit restricted to a gene range reproduces that range of the accumulators bit for bit;
* two independent algorithms agree on every gene: Step 3 bound-and-prune vs exact sort of every slice (o identical),
  Step 1 bucket selection vs warp sorters (sum of delta-IQR to 1e-9, Fisher sum within the f32-moment tolerance);
* the device ranking is a permutation that sorts the scores;
* a handful of genes against the oracle (phet_oracle.step1_all on those columns, all 1000 subsamples):
  rtol 1e-3, the float32-staging tolerance of the north star.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

CONFIGS = {"config3_10k_x_30k": (10000, 30000, 1000), "config4_50k_x_30k": (50000, 30000, 1000)}


def _run(monkeypatch, X, N, G, S, ctrl, case, idx, env=None, keep_h=False, g_range=None):
    """One fit through the C ABI; returns (scores, acc (2, G), o (G,), counts, ranking)."""
    import bench
    from phet_b200 import _capi
    for k in ("PHET_STEP1_IMPL", "PHET_STEP3_IMPL"):
        monkeypatch.delenv(k, raising=False)
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    ctx = _capi.Context(0)
    try:
        fit = bench.Fit(ctx, dict(N=N, G=G, S=S), ctrl, case, idx, 0, 1, None, perms="device")
        ctx.set_option(_capi.OPT_KEEP_H, 1 if keep_h else 0)
        ctx.set_classes(ctrl, case)
        ctx.set_index_sets(idx)
        if g_range is None:
            scores = fit.step_resident(X.data_ptr())
            order, _ = ctx.rank()
            counts = fit.counts
        else:
            ctx.run_pipeline(X.data_ptr(), N, G, _capi.PHET_F32, fit.prm, fit.m, fit.tab_full, fit.tab_last, fit.n_last,
                             g_range=g_range, s_range=(0, S), g3_range=g_range, seed=bench.SEED, is_device=True)
            scores = order = counts = None
        acc = ctx.read(_capi.BUF_ACC, np.float64, (2, G))
        o = ctx.read(_capi.BUF_O, np.float64, (G,))
        path = ctx.step3_path()
    finally:
        ctx.close()
    return scores, acc, o, counts, order, path


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_size_properties(name, monkeypatch):
    import torch
    import bench
    from oracle import phet_oracle as po
    from phet_b200 import hostmath
    N, G, S = CONFIGS[name]
    m = hostmath.subsample_size(N // 2, N - N // 2, "sqrt")
    X = bench.synth_matrix_device(N, G, torch.device("cuda", 0))
    ctrl, case, idx, _ = bench.draw_index_sets(N, S, m)

    scores, acc, o, counts, order, path = _run(monkeypatch, X, N, G, S, ctrl, case, idx)
    assert path == "bound-and-prune"
    # ---- the score vector
    assert scores.shape == (G,) and np.isfinite(scores).all() and (scores >= 0).all()
    assert abs(scores.sum() - 2.0) < 1e-9
    assert counts.sum() == G and (counts >= 0).all()
    assert np.isfinite(acc).all() and (acc[1] >= 0).all() and ((o >= 0) & (o <= 1)).all()
    k = max(G // 20, 1)   # the 5 % mean-shifted genes carry the signal (bench.synth_matrix_device)
    assert np.median(scores[:k]) > 2 * np.median(scores[2 * k:])
    # ---- ranking: a permutation that sorts the scores, ties in gene order
    assert np.array_equal(np.sort(order), np.arange(G))
    assert np.array_equal(order, np.argsort(-scores, kind="stable"))
    # ---- determinism
    scores2, acc2, o2, counts2, order2, _ = _run(monkeypatch, X, N, G, S, ctrl, case, idx)
    assert np.array_equal(scores, scores2) and np.array_equal(acc, acc2) and np.array_equal(o, o2)
    assert np.array_equal(counts, counts2) and np.array_equal(order, order2)
    # ---- gene independence: a slice of the genes on its own
    g0, g1 = G // 3, G // 3 + 1777
    _, acc_s, o_s, _, _, _ = _run(monkeypatch, X, N, G, S, ctrl, case, idx, g_range=(g0, g1))
    # (the column statistics are merged from row chunks that depend on N alone, so any partition of the genes --
    # pipeline blocks, multi-GPU shards -- sees the same z-scores)
    assert np.array_equal(o_s[g0:g1], o[g0:g1])
    assert np.array_equal(acc_s[:, g0:g1], acc[:, g0:g1])
    assert not acc_s[:, :g0].any() and not acc_s[:, g1:].any()
    # ---- two algorithms, one answer
    _, acc_w, o_w, _, _, path_w = _run(monkeypatch, X, N, G, S, ctrl, case, idx, keep_h=True,
                                       env={"PHET_STEP1_IMPL": "warp", "PHET_STEP3_IMPL": "sort"})
    assert "bound-and-prune" not in path_w
    assert np.array_equal(o_w, o)                                    # KS minimum: identical on all genes
    _, _, o_p, _, _, path_p = _run(monkeypatch, X, N, G, S, ctrl, case, idx, env={"PHET_STEP3_IMPL": "scatter"})
    assert path_p == "scatter bound-and-prune" and np.array_equal(o_p, o)    # streaming formulation of bound-and-prune: the same
    assert np.allclose(acc_w[1], acc[1], rtol=1e-9, atol=1e-9)       # sum of delta-IQR: same order statistics
    assert np.allclose(acc_w[0], acc[0], rtol=1e-4, atol=1e-6)       # Fisher sums: fp32 chunk moments vs fp64
    # ---- a few genes against the oracle, every subsample
    genes = np.array([0, 1, k, k + 1, G // 2, G - 1])
    Xc = X[:, torch.from_numpy(genes).to(X.device)].cpu().numpy().astype(np.float64)
    Xz = po.zscore_nan_to_num(Xc)
    P, R = po.step1_all(Xz, idx[:, 0].astype(np.int64), idx[:, 1].astype(np.int64), bench.PERCENTILES, m)
    fsum = po.fisher_terms(P).sum(axis=1)
    assert np.allclose(acc[1, genes], R.sum(axis=1), rtol=1e-3, atol=1e-6)
    assert np.allclose(acc[0, genes], fsum, rtol=1e-3, atol=1e-3)
