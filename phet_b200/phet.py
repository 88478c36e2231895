"""Drop-in ``PHet`` model class whose ``fit_predict`` runs on a B200 through libphet_b200.

Mirrors the reference's operator interface for this path -- constructor
``/root/reference/src/model/phet.py:33-40`` and ``fit_predict(X, y, sample_ids=None,
subtypes=None) -> ndarray (G, 1) float64`` (``phet.py:163-518``) -- with the same argument
names, defaults, error behaviour and RNG consumption, so the caller at ``src/train.py:474-477``
can switch by changing one import.  Scope is the binary (control/case), ``delta_type="iqr"``,
non-pseudobulk branch; the other branches raise ``NotImplementedError``.

There is no CPU fallback: every stage of the hot path runs in hand-written sm_100a kernels
behind the C ABI in ``include/phet_b200.h``; if the library or a B200 is missing,
``fit_predict`` raises ``phet_b200._capi.PhetError``.
"""
from __future__ import annotations

from typing import Literal, Optional

import numpy as np

from . import _capi, hostmath

SEED_VALUE = 0.001  # phet.py:29 (applied on the device, csrc/finalize.cu)

# One library context per device, kept between fits: its device buffers, streams and pinned slots are reused instead
# of being allocated and freed (cudaMalloc / cudaFree / cudaHostAlloc of ~12 GB cost ~0.1 s per fit at 50k x 30k).
_CTX_CACHE: dict = {}


def release_device_contexts():
    """Free the cached per-device contexts (device memory goes back to the driver)."""
    for ctx in list(_CTX_CACHE.values()):
        ctx.close()
    _CTX_CACHE.clear()


def _cached_context(dev: int, torch_stream: bool = False):
    """The context of device ``dev``; with ``torch_stream`` the one that launches on a (cached, non-default) torch
    stream, so that torch.distributed collectives can be ordered with the library's kernels."""
    import os
    key = (dev, bool(torch_stream))
    # the library reads its tuning / testing knobs when a context is created: a context made under other knobs is stale
    knobs = tuple(os.environ.get(k) for k in ("PHET_STEP1_IMPL", "PHET_STEP3_IMPL", "PHET_STEP1_FORCE_BIG",
                                              "PHET_STEP1_BIG_THREADS", "PHET_STEP1_NBW", "PHET_PROFILE", "PHET_EXACT_MOMENTS"))
    ctx = _CTX_CACHE.get(key)
    if ctx is not None and getattr(ctx, "knobs", None) != knobs:
        ctx.close()
        ctx = None
    if ctx is None or getattr(ctx, "h", None) is None:
        if torch_stream:
            import torch
            ts = torch.cuda.Stream(dev)            # non-default: handle 0 would mean "own stream"
            ctx = _capi.Context(dev, ts.cuda_stream)
            ctx.torch_stream = ts
        else:
            ctx = _capi.Context(dev, None)
        ctx.knobs = knobs
        if not _CTX_CACHE:
            import atexit
            atexit.register(release_device_contexts)
        _CTX_CACHE[key] = ctx
    return ctx


class PHet:
    def __init__(self, normalize: Literal["robust", "zscore", "log"] | None = "zscore",
                 percentiles: Optional[tuple[int, int]] = (25, 75), num_subsamples: int = 1000,
                 subsampling_size: Literal["optimum", "sqrt"] | int = "sqrt",
                 delta_type: Literal["iqr", "hvf"] = "iqr",
                 group_test: Literal["anova", "kruskal"] = "anova",
                 multiconditions_strategy: Literal["pairwise", "one_vs_all"] = "one_vs_all",
                 multiconditions_aggregation: Literal["mean", "max"] = "max",
                 adjust_pvalue: bool = False, feature_weight: Optional[list[float]] = None,
                 *, iqr_range: Optional[tuple[int, int]] = None, device: int | None = None,
                 storage: Literal["auto", "f32", "f64"] = "auto",
                 permutation: Literal["reference", "numpy", "device"] = "reference",
                 distributed: bool | Literal["auto"] = "auto",
                 shard: Literal["subsamples", "genes"] = "genes", capture: bool = False):
        """Same parameters as the reference (phet.py:33-104).  Extra keyword-only arguments:

        iqr_range    alias of ``percentiles`` -- it is the name ``src/train.py:474`` passes.
        device       CUDA device index (default: LOCAL_RANK if set, else 0).
        storage      element type of the normalised matrix kept in HBM/SMEM ("auto" = X's dtype).
        permutation  "reference" (default): every random draw of the fit -- the S x 2 index sets (phet.py:401-402)
                     and the 2 G Step-3 permutations (phet.py:486-487) -- is the reference's own: the
                     ``np.random`` state is handed to the library, which replays NumPy's MT19937 +
                     masked-rejection stream natively (host thread: the sequential scan; device: the swaps)
                     while the device works, and the advanced state is put back into ``np.random``.
                     Bit-compatible with the reference for a given seed.
                     "numpy": the same draws made by calling ``np.random`` in Python loops (slow: 26 s at
                     50k cells x 30k genes; kept as the cross-check of the native replay).
                     "device": index sets as the reference draws them, Step-3 permutations from a keyed
                     bijection evaluated on the GPU (no per-gene draws; NOT the reference's stream).
        distributed  shard subsamples (Step 1) and genes (Step 3) over the ranks of an initialised
                     ``torch.distributed`` group; "auto" = on when world_size > 1.
        shard        with distributed ranks: "genes" (default) = every rank loads only its gene slice
                     and runs all subsamples on it (nothing is replicated; also the layout for
                     matrices that do not fit one GPU); "subsamples" = every rank holds the whole
                     matrix and a slice of the index sets, Step 3 is gene-sliced (the north-star
                     layout; normalisation is then replicated).  Same two all-reduces either way.
        capture      keep per-stage intermediates (P, delta, H, o, bins) as attributes.
        """
        if iqr_range is not None:
            percentiles = iqr_range
        self.normalize = normalize
        self.iqr_range = percentiles
        self.num_subsamples = num_subsamples
        self.subsampling_size = subsampling_size
        self.delta_type = delta_type
        self.group_test = group_test
        self.multiconditions_strategy = multiconditions_strategy
        self.multiconditions_aggregation = multiconditions_aggregation
        self.adjust_pvalue = adjust_pvalue
        if delta_type == "hvf" and self.normalize is not None:
            self.normalize = "log"  # phet.py:116-117
        if feature_weight is None:
            feature_weight = [0.4, 0.3, 0.2, 0.1]
        elif len(feature_weight) < 2:
            raise ValueError("Feature weight list must contain at least two elements.")  # phet.py:122-123
        self.feature_weight = np.array(feature_weight) / np.sum(feature_weight)  # phet.py:124
        self.device = device
        self.storage = storage
        self.permutation = permutation
        self.distributed = distributed
        if shard not in ("subsamples", "genes"):
            raise ValueError("shard must be 'subsamples' or 'genes'")
        self.shard = shard
        self.capture = capture
        self.timings_ = {}

    @property
    def ranking_(self):
        """Gene index at each rank, descending score, equal scores in ascending gene order (phet_rank)."""
        if getattr(self, "_ranking", None) is None:
            if getattr(self, "_rank_scores", None) is None:
                raise AttributeError("ranking_ is available after fit_predict")
            ctx = _cached_context(self._rank_device)
            self._ranking, _ = ctx.rank(self._rank_scores)
        return self._ranking

    @ranking_.setter
    def ranking_(self, value):
        self._ranking = value

    # ------------------------------------------------------------------------------------
    def _check_scope(self, num_classes, sample_ids):
        if num_classes > 2:
            if self.multiconditions_strategy not in ("one_vs_all", "pairwise"):
                raise ValueError("multiconditions_strategy must be 'one_vs_all' or 'pairwise'")
            if self.multiconditions_aggregation not in ("max", "mean"):
                raise ValueError("multiconditions_aggregation must be 'max' or 'mean'")
        if sample_ids:
            raise NotImplementedError("pseudobulk branch (phet.py:224-253) is out of scope")
        if self.delta_type != "iqr":
            raise NotImplementedError("delta_type='hvf' (phet.py:255-262,403-411) is out of scope")
        if self.normalize not in ("zscore", None):
            raise NotImplementedError("normalize must be 'zscore' or None on the B200 path")
        if self.adjust_pvalue:
            raise NotImplementedError("adjust_pvalue=True (phet.py:430-431) is out of scope")

    def _resolve_device(self):
        if self.device is not None:
            return int(self.device)
        import os
        return int(os.environ.get("LOCAL_RANK", "0"))

    def _dist(self):
        """(rank, world, torch.distributed module or None)."""
        if self.distributed is False:
            return 0, 1, None
        try:
            import torch.distributed as dist
        except Exception:
            return 0, 1, None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size(), dist
        return 0, 1, None

    # ------------------------------------------------------------------------------------
    def draw_index_sets(self, examples_i, examples_j, m, replace):
        """Replay of phet.py:397-402 on the legacy global RNG: control draw then case draw."""
        S = self.num_subsamples
        idx = np.empty((S, 2, m), dtype=np.int32)
        for s in range(S):
            idx[s, 0] = np.random.choice(a=examples_i, size=m, replace=replace)
            idx[s, 1] = np.random.choice(a=examples_j, size=m, replace=replace)
        return idx

    @staticmethod
    def draw_permutations(n0, n1, num_features, min_c):
        """Replay of phet.py:486-487: per gene, permutation of the control then the case column.
        ``np.random.permutation(values)`` consumes the stream exactly like ``permutation(len)``."""
        p0 = np.empty((num_features, min_c), dtype=np.uint32)
        p1 = np.empty((num_features, min_c), dtype=np.uint32)
        for g in range(num_features):
            p0[g] = np.random.permutation(n0)[:min_c]
            p1[g] = np.random.permutation(n1)[:min_c]
        return p0, p1

    # ------------------------------------------------------------------------------------
    def fit_predict(self, X, y, sample_ids: list | None = None, subtypes: list | None = None):
        """Feature scores, shape (num_features, 1), float64 -- see phet.py:163-209."""
        import time
        t_all = time.perf_counter()
        from .ingest import StagedMatrix
        staged = X if isinstance(X, StagedMatrix) else None   # already in HBM (phet_b200.ingest.load_mtx)
        if staged is None:
            X = np.asarray(X)
        num_examples, num_features = X.shape
        y = np.asarray(y)
        unique_classes = np.unique(y)
        num_classes = len(unique_classes)
        if num_classes < 2:
            raise Exception("More than two valid groups are allowed!")  # phet.py:216-218
        if sample_ids:
            if len(sample_ids) != num_examples:
                raise Exception("The number of sample IDs provided does not match the number of example!")
            if subtypes and len(subtypes) != num_examples:
                raise Exception("The number of subtypes provided does not match the number of examples!")
        self._check_scope(num_classes, sample_ids)
        if len(y) != num_examples:
            raise Exception("The number of labels does not match the number of examples!")
        y = np.searchsorted(unique_classes, y)  # LabelEncoder, phet.py:220
        if num_classes > 2:
            from .multiclass import fit_multiclass
            scores = fit_multiclass(self, X, y, num_classes, staged)
            self.timings_["fit_s"] = time.perf_counter() - t_all
            return scores

        examples_i = np.where(y == 0)[0]
        examples_j = np.where(y == 1)[0]
        n0, n1 = len(examples_i), len(examples_j)
        min_c = min(n0, n1)
        m = hostmath.subsample_size(n0, n1, self.subsampling_size)
        if m < 1:
            raise ValueError("subsampling size must be >= 1")
        replace = bool(m > n0 or m > n1)  # phet.py:393-396
        S = int(self.num_subsamples)
        rank, world, dist = self._dist()

        if self.permutation not in ("reference", "numpy", "device"):
            raise ValueError("permutation must be 'reference', 'numpy' or 'device'")
        # ---- device and streams first: every collective below runs on this rank's own GPU
        dev = self._resolve_device()
        if staged is not None and dist is not None:
            raise NotImplementedError("a StagedMatrix is fitted by the process that staged it (single device)")
        n_slices, n_last = hostmath.step3_slices(min_c, m)
        if dist is None:
            ctx = staged.ctx if staged is not None else _cached_context(dev)
            try:
                scores = self._fit_on(ctx, X, examples_i, examples_j, m, replace, S, n_slices, n_last, 0, 1, None)
            except BaseException:
                if staged is None:          # a failed fit may leave the context half-configured: start clean next time
                    _CTX_CACHE.pop((dev, False), None)
                    ctx.close()
                raise
        else:
            import torch
            if dist.get_backend() == "nccl":
                torch.cuda.set_device(dev)                # before the first collective of this fit
            ctx = _cached_context(dev, torch_stream=True)
            with torch.cuda.stream(ctx.torch_stream):     # the caller's current stream is restored on exit
                try:
                    scores = self._fit_on(ctx, X, examples_i, examples_j, m, replace, S, n_slices, n_last, rank, world, dist)
                except BaseException:
                    _CTX_CACHE.pop((dev, True), None)
                    ctx.close()
                    raise
        self.timings_["fit_s"] = time.perf_counter() - t_all
        return scores

    # ------------------------------------------------------------------------------------
    def _fit_on(self, ctx, X, examples_i, examples_j, m, replace, S, n_slices, n_last, rank, world, dist):
        """Draws + device work on an open context.  The random draws follow the reference's order: all Step-1
        index sets, then (reference / numpy modes) the Step-3 permutations gene by gene."""
        import time
        n0, n1 = len(examples_i), len(examples_j)
        min_c = min(n0, n1)
        num_features = X.shape[1]
        t0 = time.perf_counter()
        idx = perm0 = perm1 = mt = None
        seed = 0
        draw_on_device = False
        if self.permutation == "numpy":
            if dist is None or rank == 0:
                idx = self.draw_index_sets(examples_i, examples_j, m, replace)
                perm0, perm1 = self.draw_permutations(n0, n1, num_features, min_c)
            else:
                idx = np.empty((S, 2, m), dtype=np.int32)
                perm0 = np.empty((num_features, min_c), dtype=np.uint32)
                perm1 = np.empty((num_features, min_c), dtype=np.uint32)
            if dist is not None:
                from .distributed import broadcast_numpy
                idx = broadcast_numpy(idx, dist)
                perm0 = broadcast_numpy(perm0.view(np.int32), dist).view(np.uint32)
                perm1 = broadcast_numpy(perm1.view(np.int32), dist).view(np.uint32)
        else:
            # the global stream, held natively; every rank replays rank 0's stream (no draws cross the wire)
            st = np.random.get_state()
            key, pos = np.ascontiguousarray(st[1], dtype=np.uint32), int(st[2])
            if dist is not None:
                from .distributed import broadcast_numpy
                packed = broadcast_numpy(np.concatenate([key.view(np.int32), np.array([pos], dtype=np.int32)]), dist)
                key, pos = packed[:624].view(np.uint32).copy(), int(packed[624])
            mt = _capi.MtStream(key, pos, (st[3], st[4]))
            if replace:
                # m exceeds a class (integer subsampling_size only): choice(rows, m, True) == rows[randint(0, n, m)]
                idx = np.empty((S, 2, m), dtype=np.int32)
                for s in range(S):
                    idx[s, 0] = examples_i[mt.bounded(n0 - 1, m)]
                    idx[s, 1] = examples_j[mt.bounded(n1 - 1, m)]
            else:
                draw_on_device = True
        self.timings_["host_draws_s"] = time.perf_counter() - t0
        final_state = None
        try:
            scores = self._run_device(ctx, X, examples_i, examples_j, idx, perm0, perm1, seed, m, S, n_slices, n_last,
                                      rank, world, dist, mt=mt, draw_on_device=draw_on_device)
            if dist is not None and world > 1 and draw_on_device and self.permutation == "reference":
                # every rank stopped replaying after its own genes' draws (the scan of the stream is sequential: rank r
                # pays (r + 1) / N of it); the last rank ran to the end of the stream and hands its final state over
                from .distributed import broadcast_numpy
                key, pos = mt.state()
                packed = broadcast_numpy(np.concatenate([key.view(np.int32), np.array([pos], dtype=np.int32)]), dist,
                                         src=world - 1)
                final_state = ("MT19937", packed[:624].view(np.uint32).copy(), int(packed[624]), st[3], st[4])
            return scores
        finally:
            if mt is not None:
                # np.random continues where the reference's stream would
                if final_state is not None:
                    np.random.set_state(final_state)
                else:
                    mt.to_numpy()
                mt.close()

    # ------------------------------------------------------------------------------------
    def _step1_params(self, X_dtype, m):
        qdt = np.float32 if X_dtype == np.float32 else np.float64
        (j0, k0, g0), (j1, k1, g1) = hostmath.quantile_positions(m, self.iqr_range, dtype=qdt)
        prm = _capi.Step1Params()
        prm.test = _capi.TEST_T if m < 30 else _capi.TEST_Z  # phet.py:425-428
        df = 2.0 * m - 2.0
        prm.ln_beta = hostmath.ln_beta_half(df) if (prm.test == _capi.TEST_T and df > 0) else 0.0
        # float32 input: scipy's stdtr returns float32, which underflows to 0 below 2^-150
        prm.p_flush_below = 2.0 ** -150 if (X_dtype == np.float32 and prm.test == _capi.TEST_T) else 0.0
        prm.q = _capi.QuantilePos(j0, k0, j1, k1, g0, g1)
        prm.capture_debug = 1 if self.capture else 0
        return prm

    def _run_device(self, ctx, X, examples_i, examples_j, idx, perm0, perm1, seed, m, S, n_slices, n_last,
                    rank, world, dist, mt=None, draw_on_device=False):
        N, G = X.shape
        from .ingest import StagedMatrix
        staged = isinstance(X, StagedMatrix)
        if not staged:
            if X.dtype not in (np.float32, np.float64):
                X = X.astype(np.float64)
            X = np.ascontiguousarray(X)
        dtype = _capi.PHET_F32 if X.dtype == np.float32 else _capi.PHET_F64
        storage = {"auto": None, "f32": _capi.PHET_F32, "f64": _capi.PHET_F64}[self.storage]
        from .distributed import shard_range
        if self.shard == "subsamples":
            g_range, s_range, g3_range = (0, G), shard_range(S, rank, world), shard_range(G, rank, world)
        else:
            g_range = g3_range = shard_range(G, rank, world)
            s_range = (0, S)
        ctx.set_option(_capi.OPT_KEEP_H, 1 if self.capture else 0)   # per-slice KS statistics only for capture
        ctx.set_classes(examples_i, examples_j)
        replay = self.permutation == "reference"
        draw_arg = None
        if idx is not None:
            ctx.set_index_sets(idx)
        elif replay:
            draw_arg = (S, True)                      # drawn inside the pipeline, ahead of the Step-3 draws
        else:                                         # "device": index sets from the stream, then the seed, as ever
            idx = ctx.draw_index_sets(mt, S, m)
        if self.permutation == "device":
            seed = int(mt.bounded(2 ** 31 - 2, 1)[0])  # == np.random.randint(0, 2**31 - 1)
        prm = self._step1_params(X.dtype, m)
        # Stage 0 + Step 1 + Step 3, gene-block pipelined with the H2D copy (phet.py:291-299, 397-432, 473-498)
        drawn = ctx.run_pipeline(X.device_ptr if staged else X, N, G, dtype, prm, m, hostmath.ks_exact_table(m),
                                 hostmath.ks_exact_table(n_last), n_last, is_device=staged, g_range=g_range, s_range=s_range,
                                 g3_range=g3_range,
                                 normalize=_capi.NORM_ZSCORE if self.normalize == "zscore" else _capi.NORM_NONE,
                                 storage=storage, perm_ctrl=perm0, perm_case=perm1, seed=seed, mt=mt if replay else None,
                                 replay_perms=replay, draw_index_sets=draw_arg,
                                 stop_after_own_genes=replay and dist is not None and world > 1)
        if drawn is not None:
            idx = drawn
        if replay:
            self.timings_.update(ctx.replay_stats())
        if dist is not None:
            from .distributed import allreduce_device_buffer
            # sum_s -2 ln p | sum_s delta | o, contiguous: one all-reduce (genes outside a rank's slice are 0)
            allreduce_device_buffer(ctx, _capi.BUF_ACC_O, dist)
        # Fisher standardisation, KBinsDiscretizer('uniform') on o, weights and combine (phet.py:435-462, 500-516)
        scores, counts, edges = ctx.finalize(S, self.feature_weight)
        self.bin_counts_ = counts
        # descending order of the scores (utils.py:107-136 on the device: phet_rank) is computed on first use of
        # ``ranking_``: a fit that only wants the scores does not pay for the launch and its synchronisation
        self._ranking = None
        self._rank_scores = scores
        self._rank_device = ctx.device
        self.m_ = m
        self.index_sets_ = idx
        self.launches_ = ctx.launch_count()
        self.step1_geometry_ = ctx.step1_geometry()
        self.step3_path_ = ctx.step3_path()
        self.o_ = ctx.read(_capi.BUF_O, np.float64, (G,))           # min-over-slices KS p-value per gene
        if self.capture:
            self.layout_rows_ = ctx.read(_capi.BUF_ROWS, np.int32, (N,))
            self.device_seed_ = seed
            self.fsum_, self.rsum_ = ctx.read(_capi.BUF_ACC, np.float64, (2, G))
            self.f_ = ctx.read(_capi.BUF_F, np.float64, (G,))
            self.r_ = ctx.read(_capi.BUF_R, np.float64, (G,))
            self.H_ = ctx.read(_capi.BUF_H, np.int32, (G, n_slices))
            self.bins_ = ctx.read(_capi.BUF_BINS, np.int8, (G,))
            self.edges_ = edges
            if world == 1:
                self.P_ = ctx.read(_capi.BUF_P, np.float64, (G, S))
                self.R_ = ctx.read(_capi.BUF_DELTA, np.float64, (G, S))
        return np.reshape(scores, (G, 1))
