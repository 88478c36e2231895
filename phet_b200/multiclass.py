"""Multi-class branch of ``PHet.fit_predict`` on the B200: K >= 3 conditions,
``delta_type="iqr"``, ``group_test="anova"``, both ``multiconditions_strategy`` values and both aggregations.

Mirrors ``/root/reference/src/model/phet.py`` for that branch, with the reference's RNG consumption:

* 301-306  combinations: ``one_vs_all`` = (class i, every other cell) for each i; ``pairwise`` = all class pairs;
* 309-332  subsample size over all classes;
* 348-382  Step 1a -- per subsample, one-way ANOVA over K groups of m cells  -> ``phet_anova`` (csrc/anova.cu);
* 385-417  Step 1b -- per combination and subsample, |IQR_i - IQR_j|          -> ``phet_step1`` (the two-class kernel);
* 435-444  mean over subsamples, max / mean over combinations (host, (G, C) doubles);
* 465-498  Step 3 -- per gene and combination, KS over slices whose starts cover [0, min_size) of the SMALLEST of
           all classes; the last slice runs to the end of the shorter column (it can be longer than m)
           -> ``phet_step3`` with ``PHET_OPT_STEP3_MIN_SIZE``; o = mean over combinations;
* 500-516  binning and combine                                               -> ``phet_bin`` / ``phet_combine``.

Each combination regroups the cells (class i first), so the matrix is laid out once per combination; the K ANOVA
groups are addressed through the first layout.  Labels must be 0..K-1 already (the reference mixes pre- and
post-LabelEncoder labels).  ``group_test="kruskal"`` (taken when m > 5, phet.py:372-377) -> ``phet_kruskal``.
"""
from __future__ import annotations

import math
import time
from itertools import combinations

import numpy as np

from . import _capi, hostmath


def _combos(K, strategy):
    if strategy == "one_vs_all":
        return list(zip(range(K), [0] * K))          # phet.py:304-306
    return list(combinations(range(K), 2))           # phet.py:302-303


def subsample_size_multi(y, K, subsampling_size, strategy):
    """phet.py:309-332."""
    sizes = [int(np.sum(y == c)) for c in range(K)]
    if subsampling_size == "optimum":
        out = []
        for i, j in _combos(K, strategy):
            ni = sizes[i]
            nj = (len(y) - ni) if strategy == "one_vs_all" else sizes[j]
            out.append(hostmath.optimum_sample_size(ni, nj, alpha=0.05, margin_error=0.1))
        return int(np.min(out))
    if subsampling_size == "sqrt":
        mn = int(np.min(sizes))
        if mn <= 2 or int(np.sqrt(mn)) < 8:
            return mn
        return int(np.sqrt(mn))
    return int(subsampling_size)


def plan_and_draw(est, y, K, G):
    """Combinations, sizes and every host draw of a multi-class fit, consuming ``np.random`` in the reference's
    order: S x K ANOVA draws; per combination S x (i draw, j draw); per gene and combination (i perm, j perm)."""
    strategy = est.multiconditions_strategy
    S = int(est.num_subsamples)
    combos = _combos(K, strategy)
    C = len(combos)
    rows = [np.where(y == c)[0] for c in range(K)]
    sizes = [len(r) for r in rows]
    members = []
    for i, j in combos:
        members.append((rows[i], np.where(y != i)[0] if strategy == "one_vs_all" else rows[j]))   # phet.py:389-392
    m = subsample_size_multi(y, K, est.subsampling_size, strategy)
    if m < 2:
        raise ValueError("subsampling size must be >= 2 for the multi-class branch")
    min_size = int(np.min(sizes))
    n_slices = (min_size + m - 1) // m
    start_last = (n_slices - 1) * m
    replace_a = any(m > n for n in sizes)                                   # phet.py:359-360
    idx_anova = np.empty((S, K, m), dtype=np.int32)
    for s in range(S):
        for c in range(K):
            t = np.random.choice(a=sizes[c], size=m, replace=replace_a)     # == choice(range(n_c), ...) phet.py:369
            idx_anova[s, c] = rows[c][t]
    idx_combo = []
    for ei, ej in members:
        rep = bool(m > len(ei) or m > len(ej))                              # phet.py:395-396
        idx = np.empty((S, 2, m), dtype=np.int32)
        for s in range(S):
            idx[s, 0] = np.random.choice(a=ei, size=m, replace=rep)
            idx[s, 1] = np.random.choice(a=ej, size=m, replace=rep)
        idx_combo.append(idx)
    n_last = [min(len(ei), len(ej)) - start_last for ei, ej in members]     # phet.py:489-491
    perms = None
    seed = 0
    if est.permutation == "reference":
        perms = [(np.empty((G, start_last + n_last[c]), dtype=np.uint32), np.empty((G, start_last + n_last[c]), dtype=np.uint32))
                 for c in range(C)]
        for g in range(G):                                                  # phet.py:473-487: gene-major, then combination
            for c, (ei, ej) in enumerate(members):
                perms[c][0][g] = np.random.permutation(len(ei))[:start_last + n_last[c]]
                perms[c][1][g] = np.random.permutation(len(ej))[:start_last + n_last[c]]
    elif est.permutation == "device":
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    else:
        raise ValueError("permutation must be 'reference' or 'device'")
    return dict(combos=combos, members=members, m=m, min_size=min_size, n_slices=n_slices, n_last=n_last,
                idx_anova=idx_anova, idx_combo=idx_combo, perms=perms, seed=seed)


def fit_multiclass(est, X, y, K, staged):
    """Scores (G, 1) float64 for K >= 3 classes.  ``est``: the PHet instance (parameters in, attributes out)."""
    rank, world, dist = est._dist()
    if dist is not None and staged is not None:
        raise NotImplementedError("a StagedMatrix is fitted by the process that staged it (single device)")
    if est.group_test not in ("anova", "kruskal"):
        raise ValueError("group_test must be 'anova' or 'kruskal'")
    if est.multiconditions_strategy == "pairwise" and est.permutation == "device":
        raise NotImplementedError("pairwise combinations need permutation='reference' (the device walk covers a whole group)")
    agg = est.multiconditions_aggregation
    N, G = X.shape
    S = int(est.num_subsamples)
    if dist is not None and dist.get_backend() == "nccl":
        import torch
        torch.cuda.set_device(est._resolve_device())   # before the first collective: broadcast_numpy allocates on the current device
    t0 = time.perf_counter()
    if dist is None or rank == 0:
        plan = plan_and_draw(est, y, K, G)
    if dist is not None:
        # only rank 0's NumPy stream matters (as in the two-class path): its draws are broadcast
        from .distributed import broadcast_numpy
        import pickle
        blob = np.frombuffer(pickle.dumps(plan), dtype=np.uint8) if rank == 0 else np.zeros(0, dtype=np.uint8)
        size = int(broadcast_numpy(np.array([blob.size], dtype=np.int64), dist)[0])
        if rank != 0:
            blob = np.zeros(size, dtype=np.uint8)
        plan = pickle.loads(broadcast_numpy(blob, dist).tobytes())
    members, m, min_size, n_slices, n_last = plan["members"], plan["m"], plan["min_size"], plan["n_slices"], plan["n_last"]
    idx_anova, idx_combo, perms, seed = plan["idx_anova"], plan["idx_combo"], plan["perms"], plan["seed"]
    C = len(members)
    est.timings_["host_draws_s"] = time.perf_counter() - t0

    # ---- device
    is_staged = staged is not None
    if not is_staged:
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        X = np.ascontiguousarray(X)
    dtype = _capi.PHET_F32 if X.dtype == np.float32 else _capi.PHET_F64
    storage = {"auto": dtype, "f32": _capi.PHET_F32, "f64": _capi.PHET_F64}[est.storage]
    norm = _capi.NORM_ZSCORE if est.normalize == "zscore" else _capi.NORM_NONE
    prm = est._step1_params(X.dtype, m)
    prm.capture_debug = 0
    a, b = 0.5 * (K * m - K), 0.5 * (K - 1)
    ln_beta = math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
    p_flush = 2.0 ** -150 if X.dtype == np.float32 else 0.0                  # float32 fdtrc underflows to 0
    ctx = staged.ctx if is_staged else _capi.Context(est._resolve_device())
    rsum = np.empty((G, C))
    o_combo = np.empty((G, C))
    H = []
    try:
        ctx.set_option(_capi.OPT_KEEP_H, 1 if est.capture else 0)
        # gene sharding (the capacity layout): a rank lays out, and computes, only its gene slice; gene
        # numbering stays global, so the accumulators and the keyed permutations are those of the one-GPU fit
        from .distributed import shard_range
        g_lo, g_hi = shard_range(G, rank, world) if dist is not None else (0, G)
        if g_hi <= g_lo:
            raise ValueError("gene sharding needs at least one gene per rank")
        if is_staged:
            raw_ptr, ldx = X.device_ptr, G
        elif dist is None:
            raw_ptr, ldx = ctx.stage_matrix(X), G                            # the matrix crosses PCIe once
        else:
            raw_ptr, ldx = ctx.stage_matrix(np.ascontiguousarray(X[:, g_lo:g_hi])), g_hi - g_lo
        for c, (ei, ej) in enumerate(members):
            ctx.set_option(_capi.OPT_STEP3_MIN_SIZE, 0)
            # pairwise: the cells of the other classes ride behind the second group (the layout must hold every
            # cell: the z-score is over all of them and the ANOVA draws address them); nothing ever indexes them
            rest = np.setdiff1d(np.arange(N), np.concatenate([ei, ej]))
            ctx.set_classes(ei, np.concatenate([ej, rest]) if rest.size else ej)
            if dist is None:
                ctx.load_matrix_device(raw_ptr, N, G, dtype, norm, storage)
            else:
                ctx.load_matrix_genes(raw_ptr, ldx, N, G, dtype, g_lo, g_hi, norm, storage)
            if c == 0:
                if est.group_test == "kruskal" and m > 5:                    # phet.py:372-377
                    ctx.kruskal(idx_anova, p_flush, capture=est.capture)
                else:
                    ctx.anova(idx_anova, ln_beta, p_flush, capture=est.capture)
                fsum = ctx.read(_capi.BUF_ACC, np.float64, (2, G))[0].copy()
                if est.capture:
                    est.P_ = ctx.read(_capi.BUF_P, np.float64, (G, S))
            ctx.set_index_sets(idx_combo[c])
            ctx.step1(0, S, prm)
            rsum[:, c] = ctx.read(_capi.BUF_ACC, np.float64, (2, G))[1]
            ctx.set_option(_capi.OPT_STEP3_MIN_SIZE, min_size)
            ctx.step3(g_lo, g_hi, m, hostmath.ks_exact_table(m), hostmath.ks_exact_table(n_last[c]), n_last[c],
                      perm_ctrl=perms[c][0] if perms is not None else None,
                      perm_case=perms[c][1] if perms is not None else None, seed=seed + c)
            o_combo[:, c] = ctx.read(_capi.BUF_O, np.float64, (G,))
            if est.capture:
                H.append(ctx.read(_capi.BUF_H, np.int32, (G, n_slices)))
        ctx.set_option(_capi.OPT_STEP3_MIN_SIZE, 0)
        if dist is not None:
            # ranks hold zeros outside their gene slice: SUM completes the vectors exactly (one small all-reduce)
            from .distributed import allreduce_numpy
            packed = allreduce_numpy(np.concatenate([fsum[:, None], rsum, o_combo], axis=1), dist)
            fsum, rsum, o_combo = packed[:, 0].copy(), packed[:, 1:1 + C].copy(), packed[:, 1 + C:].copy()
            if est.capture:
                H = [allreduce_numpy(h, dist) for h in H]
                est.P_ = allreduce_numpy(est.P_, dist)
        # phet.py:435-444
        r = rsum / S
        r = np.mean(r, axis=1) if agg == "mean" else np.max(r, axis=1)
        np.nan_to_num(r, copy=False, nan=0.0, posinf=0.0, neginf=0.0)
        o = np.mean(o_combo, axis=1)                                         # phet.py:498
        # Fisher, binning and combine on the device (phet.py:449-462, 500-516)
        ctx.write(_capi.BUF_ACC, np.stack([fsum, np.zeros(G)]))
        ctx.fisher(S)
        ctx.write(_capi.BUF_R, r)
        ctx.write(_capi.BUF_O, o)
        lo, hi = ctx.o_minmax()
        if lo == hi:
            edges = np.array([-np.inf, np.inf])
            counts1 = ctx.bin(edges)
            counts = np.zeros(len(est.feature_weight), dtype=np.int64)
            counts[0] = counts1[0]
            scores = ctx.combine(est.feature_weight[:1])
        else:
            edges = hostmath.uniform_bin_edges(lo, hi, len(est.feature_weight))
            counts = ctx.bin(edges)
            scores = ctx.combine(est.feature_weight)
        est.ranking_, _ = ctx.rank()
        est.bin_counts_ = counts
        est.m_ = m
        est.o_ = o
        est.launches_ = ctx.launch_count()
        est.step3_path_ = ctx.step3_path()
        if est.capture:
            est.fsum_, est.rsum_combo_, est.o_combo_, est.H_combo_ = fsum, rsum, o_combo, H
            est.r_ = r
            est.f_ = ctx.read(_capi.BUF_F, np.float64, (G,))
            est.bins_ = ctx.read(_capi.BUF_BINS, np.int8, (G,))
            est.edges_ = edges
            est.index_sets_anova_, est.index_sets_combo_ = idx_anova, idx_combo
    finally:
        if not is_staged:
            ctx.close()
    return np.reshape(scores, (G, 1))
