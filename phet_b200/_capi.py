"""ctypes binding of libphet_b200.so (include/phet_b200.h).  No torch types cross this boundary.

The library is built in-tree by ``phet_b200.build`` (nvcc, sm_100a).  There is no CPU fallback:
if the library cannot be loaded, or no sm_100 device is present, the calls raise ``PhetError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libphet_b200.so")

PHET_F32, PHET_F64 = 0, 1
NORM_NONE, NORM_ZSCORE = 0, 1
TEST_T, TEST_Z = 0, 1
PERM_HOST, PERM_DEVICE, PERM_REPLAY, PERM_RESIDENT = 0, 1, 2, 3
OPT_KEEP_H = 1
OPT_STEP3_MIN_SIZE = 2
(BUF_ACC, BUF_O, BUF_H, BUF_P, BUF_DELTA, BUF_F, BUF_R, BUF_BINS, BUF_SCORES, BUF_XG, BUF_MEAN,
 BUF_STD, BUF_ROWS, BUF_PROFILE, BUF_ACC_O, BUF_STAGED, BUF_RANK_ORDER, BUF_RAW) = range(18)

EXPORTS = [
    "phet_abi_version", "phet_last_error", "phet_device_count", "phet_ctx_create", "phet_ctx_destroy",
    "phet_ctx_sync", "phet_set_classes", "phet_load_matrix", "phet_set_index_sets", "phet_step1",
    "phet_step3", "phet_run_pipeline", "phet_fisher", "phet_o_minmax", "phet_bin", "phet_combine", "phet_finalize", "phet_finalize_async", "phet_fetch_results", "phet_device_ptr",
    "phet_read_buffer", "phet_write_buffer", "phet_set_option", "phet_step3_path", "phet_launch_count", "phet_step1_geometry", "phet_fit_host", "phet_class_stats",
    "phet_mtx_parse", "phet_mtx_triplets", "phet_mtx_free", "phet_mtx_stage", "phet_mtx_kept", "phet_rank", "phet_anova", "phet_kruskal", "phet_stage_matrix", "phet_load_matrix_genes",
    "phet_host_p_two_sided_t", "phet_host_p_two_sided_z", "phet_host_perm_apply",
    "phet_mt_create", "phet_mt_state", "phet_mt_destroy", "phet_mt_simd", "phet_mt_shuffle_rounds", "phet_mt_bounded",
    "phet_mt_par_begin", "phet_mt_par_threads", "phet_mt_par_submit", "phet_mt_par_wait", "phet_mt_par_end",
    "phet_host_fisher_yates", "phet_fy_apply", "phet_replay_stats", "phet_draw_index_sets", "phet_phase_ms",
]


class PhetError(RuntimeError):
    pass


class QuantilePos(C.Structure):
    _fields_ = [("j_lo", C.c_int32), ("k_lo", C.c_int32), ("j_hi", C.c_int32), ("k_hi", C.c_int32),
                ("g_lo", C.c_double), ("g_hi", C.c_double)]


class Step1Params(C.Structure):
    _fields_ = [("test", C.c_int32), ("ln_beta", C.c_double), ("p_flush_below", C.c_double),
                ("q", QuantilePos), ("capture_debug", C.c_int32)]


class PipelineParams(C.Structure):
    _fields_ = [("normalize", C.c_int32), ("storage", C.c_int32), ("g_begin", C.c_int64), ("g_end", C.c_int64),
                ("s_begin", C.c_int64), ("s_end", C.c_int64), ("g3_begin", C.c_int64), ("g3_end", C.c_int64),
                ("step1", Step1Params), ("m", C.c_int64), ("perm_mode", C.c_int32), ("seed", C.c_uint64),
                ("perm_ctrl", C.c_void_p), ("perm_case", C.c_void_p), ("table_full", C.c_void_p),
                ("table_last", C.c_void_p), ("n_last", C.c_int64), ("block_genes", C.c_int64),
                ("mt", C.c_void_p), ("draw_index_sets", C.c_int32), ("S_draw", C.c_int64), ("idx_out", C.c_void_p),
                ("keep_perm_tables", C.c_int32), ("stop_after_own_genes", C.c_int32)]


class MtJob(C.Structure):
    _fields_ = [("n", C.c_int64), ("out", C.c_void_p), ("stride_bytes", C.c_int64), ("elem_bytes", C.c_int32)]


class FitParams(C.Structure):
    _fields_ = [("normalize", C.c_int32), ("storage", C.c_int32), ("step1", Step1Params), ("m", C.c_int64),
                ("perm_mode", C.c_int32), ("seed", C.c_uint64), ("n_bins", C.c_int32),
                ("weights", C.POINTER(C.c_double)), ("table_full", C.POINTER(C.c_double)),
                ("table_last", C.POINTER(C.c_double)), ("n_last", C.c_int64)]


_lib = None


def load_library(build_if_missing: bool = True):
    """Load libphet_b200.so (building it with nvcc first if it is absent or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        try:
            if _build.needs_build():
                _build.build()
        except Exception as exc:  # no nvcc on the box: use the shipped .so if there is one
            if not os.path.exists(LIB_PATH):
                raise PhetError("libphet_b200.so is missing and could not be built: %s" % exc)
    if not os.path.exists(LIB_PATH):
        raise PhetError("libphet_b200.so not found at %s (run `python -m phet_b200.build`); "
                        "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    lib.phet_abi_version.restype = C.c_int
    lib.phet_last_error.restype = C.c_char_p
    lib.phet_device_count.restype = C.c_int
    lib.phet_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.phet_ctx_destroy.argtypes = [vp]
    lib.phet_ctx_sync.argtypes = [vp]
    lib.phet_set_classes.argtypes = [vp, vp, i64, vp, i64]
    lib.phet_load_matrix.argtypes = [vp, vp, C.c_int, i64, i64, C.c_int, C.c_int, C.c_int]
    lib.phet_set_index_sets.argtypes = [vp, vp, i64, i64]
    lib.phet_step1.argtypes = [vp, i64, i64, C.POINTER(Step1Params), C.c_int]
    lib.phet_step3.argtypes = [vp, i64, i64, i64, C.c_int, vp, vp, u64, vp, vp, i64]
    lib.phet_run_pipeline.argtypes = [vp, vp, C.c_int, i64, i64, C.c_int, C.POINTER(PipelineParams)]
    lib.phet_fisher.argtypes = [vp, i64]
    lib.phet_o_minmax.argtypes = [vp, C.POINTER(dbl)]
    lib.phet_bin.argtypes = [vp, vp, i32, vp]
    lib.phet_combine.argtypes = [vp, vp, i32, vp]
    lib.phet_finalize.argtypes = [vp, i64, vp, i32, vp, vp, vp]
    lib.phet_finalize_async.argtypes = [vp, i64, vp, i32]
    lib.phet_fetch_results.argtypes = [vp, vp, vp, vp]
    lib.phet_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    lib.phet_read_buffer.argtypes = [vp, C.c_int, vp, i64]
    lib.phet_write_buffer.argtypes = [vp, C.c_int, vp, i64]
    lib.phet_set_option.argtypes = [vp, C.c_int, i64]
    lib.phet_step3_path.argtypes = [vp]
    lib.phet_launch_count.argtypes = [vp]
    lib.phet_launch_count.restype = i64
    lib.phet_step1_geometry.argtypes = [vp, C.POINTER(i32)]
    lib.phet_fit_host.argtypes = [vp, vp, C.c_int, i64, i64, vp, i64, vp, i64, vp, i64, vp, vp,
                                  C.POINTER(FitParams), vp, vp]
    lib.phet_class_stats.argtypes = [vp, C.POINTER(QuantilePos), vp]
    lib.phet_mtx_parse.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    lib.phet_mtx_triplets.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_int)]
    lib.phet_mtx_free.argtypes = [vp]
    lib.phet_mtx_stage.argtypes = [vp, vp, C.c_int, i64, C.POINTER(i64)]
    lib.phet_mtx_kept.argtypes = [vp, vp, vp]
    lib.phet_rank.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.phet_load_matrix_genes.argtypes = [vp, vp, i64, i64, i64, C.c_int, C.c_int, C.c_int, i64, i64]
    lib.phet_stage_matrix.argtypes = [vp, vp, i64, i64, C.c_int]
    lib.phet_kruskal.argtypes = [vp, vp, i64, i64, i64, dbl, C.c_int]
    lib.phet_anova.argtypes = [vp, vp, i64, i64, i64, dbl, dbl, C.c_int]
    lib.phet_host_p_two_sided_t.argtypes = [dbl, dbl, dbl]
    lib.phet_host_p_two_sided_t.restype = dbl
    lib.phet_host_p_two_sided_z.argtypes = [dbl]
    lib.phet_host_p_two_sided_z.restype = dbl
    lib.phet_host_perm_apply.argtypes = [u64, u64, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.phet_host_perm_apply.restype = C.c_uint32
    lib.phet_mt_create.argtypes = [vp, i32, C.POINTER(vp)]
    lib.phet_mt_state.argtypes = [vp, vp, C.POINTER(i32)]
    lib.phet_mt_destroy.argtypes = [vp]
    lib.phet_mt_simd.argtypes = [vp]
    lib.phet_mt_shuffle_rounds.argtypes = [vp, C.POINTER(MtJob), i32, i64]
    lib.phet_mt_bounded.argtypes = [vp, C.c_uint32, i64, vp]
    lib.phet_mt_par_begin.argtypes = [vp, i32, i64, C.POINTER(vp)]
    lib.phet_mt_par_threads.argtypes = [vp, C.POINTER(i32)]
    lib.phet_mt_par_submit.argtypes = [vp, C.POINTER(MtJob), i32, i64, C.POINTER(i64)]
    lib.phet_mt_par_wait.argtypes = [vp, i64]
    lib.phet_mt_par_end.argtypes = [vp]
    lib.phet_host_fisher_yates.argtypes = [vp, i32, i64, i64, vp]
    lib.phet_fy_apply.argtypes = [vp, vp, i32, i64, i64, i64, i64, vp]
    lib.phet_replay_stats.argtypes = [vp, C.POINTER(dbl)]
    lib.phet_draw_index_sets.argtypes = [vp, vp, i64, i64, vp]
    lib.phet_phase_ms.argtypes = [vp, C.POINTER(C.c_float)]
    for name in EXPORTS:
        getattr(lib, name)
    if lib.phet_abi_version() != 1:
        raise PhetError("libphet_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise PhetError("phet_b200: %s (status %d)" % (load_library().phet_last_error().decode(), rc))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class MtStream:
    """NumPy's legacy global stream (``np.random``: MT19937) held natively -- include/phet_b200.h ``phet_mt``.

    ``MtStream.from_numpy()`` takes ``np.random.get_state()``; the draws the library makes advance it exactly as
    the reference's ``np.random.choice`` / ``np.random.permutation`` calls would (phet.py:401-402, 486-487);
    ``to_numpy()`` puts the advanced state back with ``np.random.set_state`` so the caller's stream goes on where
    the reference's would.  No GPU involved."""

    def __init__(self, key: np.ndarray, pos: int, gauss=(0, 0.0)):
        self.lib = load_library()
        key = np.ascontiguousarray(key, dtype=np.uint32)
        assert key.size == 624
        h = C.c_void_p()
        _check(self.lib.phet_mt_create(_ptr(key), int(pos), C.byref(h)))
        self.h = h
        self._gauss = gauss

    @classmethod
    def from_numpy(cls, state=None):
        st = np.random.get_state() if state is None else state
        if st[0] != "MT19937":
            raise PhetError("np.random is not the legacy MT19937 stream")
        return cls(st[1], st[2], (st[3], st[4]))

    def state(self):
        key = np.empty(624, dtype=np.uint32)
        pos = C.c_int32()
        _check(self.lib.phet_mt_state(self.h, _ptr(key), C.byref(pos)))
        return key, int(pos.value)

    def numpy_state(self):
        key, pos = self.state()
        return ("MT19937", key, pos, self._gauss[0], self._gauss[1])

    def to_numpy(self):
        np.random.set_state(self.numpy_state())

    def simd(self) -> int:
        return int(self.lib.phet_mt_simd(self.h))

    def shuffle_draws(self, sizes, rounds: int, store: bool = True, workers: int | None = None, chunks: int = 1):
        """The swap targets of ``rounds`` x [np.random.permutation(n) for n in sizes]: a list of arrays
        (rounds, n - 1), uint16 when n <= 65536 else uint32 (None when ``store`` is False: stream advanced only).
        ``workers`` not None: through the multi-threaded replay (generator / scanner / store workers), the
        rounds submitted in ``chunks`` blocks (``self.par_threads``: generator, scanner, worker threads used)."""
        outs, jobs = [], (MtJob * len(sizes))()
        for k, n in enumerate(sizes):
            eb = 2 if n <= 65536 else 4
            arr = np.zeros((rounds, max(int(n) - 1, 1)), dtype=np.uint16 if eb == 2 else np.uint32) if store else None
            outs.append(arr)
            jobs[k] = MtJob(int(n), arr.ctypes.data if store else None, arr.strides[0] if store else 0, eb)
        if workers is None:
            _check(self.lib.phet_mt_shuffle_rounds(self.h, jobs, len(sizes), int(rounds)))
            return outs
        par = C.c_void_p()
        _check(self.lib.phet_mt_par_begin(self.h, int(workers), int(max(sizes)), C.byref(par)))
        try:
            tickets, r0 = [], 0
            per = max(1, (int(rounds) + chunks - 1) // chunks)
            while r0 < rounds:
                rc = min(per, rounds - r0)
                jj = (MtJob * len(sizes))()
                for k, n in enumerate(sizes):
                    jj[k] = MtJob(jobs[k].n, (jobs[k].out + r0 * jobs[k].stride_bytes) if store else None,
                                  jobs[k].stride_bytes, jobs[k].elem_bytes)
                t = C.c_int64()
                _check(self.lib.phet_mt_par_submit(par, jj, len(sizes), rc, C.byref(t)))
                tickets.append(t.value)
                r0 += rc
            for t in tickets:
                _check(self.lib.phet_mt_par_wait(par, t))
            st = (C.c_int32 * 3)()
            _check(self.lib.phet_mt_par_threads(par, st))
            self.par_threads = [int(v) for v in st]
        finally:
            _check(self.lib.phet_mt_par_end(par))
        return outs

    def permutation(self, n: int, take: int | None = None) -> np.ndarray:
        """np.random.permutation(n)[:take], swaps applied on the host (tests, small problems)."""
        n = int(n)
        take = n if take is None else int(take)
        draws = self.shuffle_draws([n], 1)[0]
        out = np.empty(take, dtype=np.uint32)
        _check(self.lib.phet_host_fisher_yates(_ptr(draws), draws.dtype.itemsize, n, take, _ptr(out)))
        return out

    def bounded(self, maximum: int, count: int) -> np.ndarray:
        """``count`` draws of rk_interval(maximum): what np.random.randint(0, maximum + 1, count) returns."""
        out = np.empty(int(count), dtype=np.uint32)
        _check(self.lib.phet_mt_bounded(self.h, int(maximum), int(count), _ptr(out)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.phet_mt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MtxFile:
    """A MatrixMarket coordinate file parsed by the library's multi-threaded host reader (phet_mtx_parse;
    replaces sc.read_mtx at src/train.py:82).  ``shape`` = (cells, genes); no GPU involved."""

    def __init__(self, path: str, threads: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        dims = (C.c_int64 * 3)()
        _check(self.lib.phet_mtx_parse(os.fsencode(path), int(threads), C.byref(h), dims))
        self.h = h
        self.shape = (int(dims[0]), int(dims[1]))
        self.entries = int(dims[2])

    def triplets(self):
        """(rows, cols, vals, symmetric): copies of the zero-based triplets in file order."""
        r, c, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n, sym = C.c_int64(), C.c_int()
        _check(self.lib.phet_mtx_triplets(self.h, C.byref(r), C.byref(c), C.byref(v), C.byref(n), C.byref(sym)))
        k = n.value
        if k == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), bool(sym.value)
        rows = np.ctypeslib.as_array(C.cast(r, C.POINTER(C.c_int32)), (k,)).copy()
        cols = np.ctypeslib.as_array(C.cast(c, C.POINTER(C.c_int32)), (k,)).copy()
        vals = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_float)), (k,)).copy()
        return rows, cols, vals, bool(sym.value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.phet_mtx_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One device, one stream.  Thin object wrapper over the ``phet_ctx`` C handle."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.phet_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device
        self.G = 0
        self.N = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.phet_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(self.lib.phet_ctx_sync(self.h))

    def set_option(self, option: int, value: int):
        _check(self.lib.phet_set_option(self.h, int(option), int(value)))

    def set_classes(self, ctrl_rows, case_rows):
        a = np.ascontiguousarray(ctrl_rows, dtype=np.int32)
        b = np.ascontiguousarray(case_rows, dtype=np.int32)
        _check(self.lib.phet_set_classes(self.h, _ptr(a), a.size, _ptr(b), b.size))
        self.n0, self.n1 = a.size, b.size

    def load_matrix(self, X: np.ndarray, normalize=NORM_ZSCORE, storage=None):
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        X = np.ascontiguousarray(X)
        dt = PHET_F32 if X.dtype == np.float32 else PHET_F64
        st = dt if storage is None else storage
        _check(self.lib.phet_load_matrix(self.h, _ptr(X), 0, X.shape[0], X.shape[1], dt, normalize, st))
        self.N, self.G = X.shape

    def stage_matrix(self, X: np.ndarray) -> int:
        """Copy a C-contiguous float32 / float64 host matrix to HBM once (BUF_RAW); returns the device pointer."""
        assert X.flags.c_contiguous and X.dtype in (np.float32, np.float64)
        _check(self.lib.phet_stage_matrix(self.h, _ptr(X), X.shape[0], X.shape[1], PHET_F32 if X.dtype == np.float32 else PHET_F64))
        return self.device_ptr(BUF_RAW)[0]

    def load_matrix_device(self, dev_ptr: int, N: int, G: int, dtype: int, normalize=NORM_ZSCORE, storage=None):
        st = dtype if storage is None else storage
        _check(self.lib.phet_load_matrix(self.h, C.c_void_p(dev_ptr), 1, N, G, dtype, normalize, st))
        self.N, self.G = N, G

    def load_matrix_genes(self, dev_ptr: int, ldx: int, N: int, G: int, dtype: int, g_begin: int, g_end: int,
                          normalize=NORM_ZSCORE, storage=None):
        """Lay out genes [g_begin, g_end) only; dev_ptr addresses element (0, g_begin), leading dimension ldx."""
        st = dtype if storage is None else storage
        _check(self.lib.phet_load_matrix_genes(self.h, C.c_void_p(dev_ptr), ldx, N, G, dtype, normalize, st, g_begin, g_end))
        self.N, self.G = N, G

    def class_stats(self, q3):
        """q3: three QuantilePos (class 0, class 1, all cells).  Returns float64 (7, G): iqr0, iqr1, iqr_all,
        mean0, mean1, ss0, ss1 (include/phet_b200.h: phet_class_stats)."""
        arr = (QuantilePos * 3)(*q3)
        out = np.empty((7, self.G), dtype=np.float64)
        _check(self.lib.phet_class_stats(self.h, arr, _ptr(out)))
        return out

    def stage_mtx(self, mtx: "MtxFile", do_filter: bool = True, minimum_samples: int = 5):
        """Dense float32 matrix of ``mtx`` in HBM, filtered like src/train.py:88-118 when ``do_filter``.
        Returns (cells kept, genes kept, kept row numbers, kept column numbers); the matrix is BUF_STAGED."""
        dims = (C.c_int64 * 2)()
        _check(self.lib.phet_mtx_stage(self.h, mtx.h, 1 if do_filter else 0, int(minimum_samples), dims))
        n, g = int(dims[0]), int(dims[1])
        rows = np.empty(n, dtype=np.int32)
        cols = np.empty(g, dtype=np.int32)
        _check(self.lib.phet_mtx_kept(self.h, _ptr(rows), _ptr(cols)))
        return n, g, rows, cols

    def rank(self, scores=None, k=None):
        """Descending order of the scores (host vector, or the resident BUF_SCORES when None):
        (gene index per rank int32[k], score per rank float64[k]) -- include/phet_b200.h: phet_rank."""
        if scores is None:
            G, sp, keep = self.G, None, None
        else:
            keep = np.ascontiguousarray(np.asarray(scores, dtype=np.float64).reshape(-1))
            G, sp = keep.size, _ptr(keep)
        k = G if k is None else int(k)
        order = np.empty(k, dtype=np.int32)
        srt = np.empty(k, dtype=np.float64)
        _check(self.lib.phet_rank(self.h, sp, G, k, _ptr(order), _ptr(srt)))
        return order, srt

    def anova(self, idx: np.ndarray, ln_beta: float, p_flush_below: float = 0.0, capture: bool = False):
        """idx: int32 (S, K, m) original rows.  Fisher's sum of the one-way ANOVA p-values -> BUF_ACC[0:G]."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        S, K, m = idx.shape
        _check(self.lib.phet_anova(self.h, _ptr(idx), S, K, m, float(ln_beta), float(p_flush_below), 1 if capture else 0))

    def kruskal(self, idx: np.ndarray, p_flush_below: float = 0.0, capture: bool = False):
        """idx: int32 (S, K, m) original rows.  Fisher's sum of the Kruskal-Wallis p-values -> BUF_ACC[0:G]."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        S, K, m = idx.shape
        _check(self.lib.phet_kruskal(self.h, _ptr(idx), S, K, m, float(p_flush_below), 1 if capture else 0))

    def set_index_sets(self, idx: np.ndarray):
        """idx: int32 (S, 2, m), original row numbers; [:, 0] control draw, [:, 1] case draw."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        assert idx.ndim == 3 and idx.shape[1] == 2
        _check(self.lib.phet_set_index_sets(self.h, _ptr(idx), idx.shape[0], idx.shape[2]))
        self.S, self.m = idx.shape[0], idx.shape[2]

    def step1(self, s_begin, s_end, params: Step1Params, accumulate=False):
        _check(self.lib.phet_step1(self.h, s_begin, s_end, C.byref(params), 1 if accumulate else 0))

    def step3(self, g_begin, g_end, m, table_full, table_last, n_last, perm_ctrl=None, perm_case=None,
              seed=0, resident_perms=False):
        tf = np.ascontiguousarray(table_full, dtype=np.float64)
        tl = np.ascontiguousarray(table_last, dtype=np.float64)
        if perm_ctrl is not None:
            pc = np.ascontiguousarray(perm_ctrl, dtype=np.uint32)
            pk = np.ascontiguousarray(perm_case, dtype=np.uint32)
            _check(self.lib.phet_step3(self.h, g_begin, g_end, m, PERM_HOST, _ptr(pc), _ptr(pk), 0, _ptr(tf),
                                       _ptr(tl), n_last))
        elif resident_perms:
            _check(self.lib.phet_step3(self.h, g_begin, g_end, m, PERM_RESIDENT, None, None, 0, _ptr(tf), _ptr(tl), n_last))
        else:
            _check(self.lib.phet_step3(self.h, g_begin, g_end, m, PERM_DEVICE, None, None, seed, _ptr(tf),
                                       _ptr(tl), n_last))

    def run_pipeline(self, X, N, G, dtype, params: Step1Params, m, table_full, table_last, n_last,
                     g_range=None, s_range=None, g3_range=None, normalize=NORM_ZSCORE, storage=None,
                     perm_ctrl=None, perm_case=None, seed=0, is_device=False, block_genes=0, mt=None,
                     replay_perms=False, draw_index_sets=None, keep_perm_tables=False, resident_perms=False,
                     stop_after_own_genes=False):
        """Stage 0 + Step 1 + Step 3 over gene blocks.  X: C-contiguous host ndarray (N, G) or, with
        is_device=True, a raw device pointer (int).  ``mt`` (an MtStream) + ``replay_perms=True``: the Step-3
        permutations are the reference's own draws (PHET_PERM_REPLAY); ``draw_index_sets=(S, keep)``: the index
        sets are drawn from ``mt`` first and, with keep=True, returned as int32 (S, 2, m) original rows."""
        pp = PipelineParams()
        idx_out = None
        if draw_index_sets is not None:
            S_draw, keep_idx = draw_index_sets
            pp.draw_index_sets, pp.S_draw = 1, int(S_draw)
            self.S, self.m = int(S_draw), int(m)
            if keep_idx:
                idx_out = np.empty((int(S_draw), 2, int(m)), dtype=np.int32)
                pp.idx_out = idx_out.ctypes.data
        if mt is not None:
            pp.mt = mt.h
        pp.normalize = normalize
        pp.storage = dtype if storage is None else storage
        pp.g_begin, pp.g_end = g_range if g_range is not None else (0, G)
        pp.s_begin, pp.s_end = s_range if s_range is not None else (0, self.S)
        pp.g3_begin, pp.g3_end = g3_range if g3_range is not None else (pp.g_begin, pp.g_end)
        pp.step1 = params
        pp.m = m
        tf = np.ascontiguousarray(table_full, dtype=np.float64)
        tl = np.ascontiguousarray(table_last, dtype=np.float64)
        pp.table_full, pp.table_last, pp.n_last = tf.ctypes.data, tl.ctypes.data, n_last
        keep = [tf, tl]
        if perm_ctrl is not None:
            pc = np.ascontiguousarray(perm_ctrl, dtype=np.uint32)
            pk = np.ascontiguousarray(perm_case, dtype=np.uint32)
            keep += [pc, pk]
            pp.perm_mode, pp.perm_ctrl, pp.perm_case = PERM_HOST, pc.ctypes.data, pk.ctypes.data
        elif replay_perms:
            pp.perm_mode = PERM_REPLAY
            pp.keep_perm_tables = 1 if keep_perm_tables else 0
            pp.stop_after_own_genes = 1 if stop_after_own_genes else 0
        elif resident_perms:
            pp.perm_mode = PERM_RESIDENT
        else:
            pp.perm_mode, pp.seed = PERM_DEVICE, seed
        pp.block_genes = block_genes
        ptr = C.c_void_p(X) if is_device else _ptr(X)
        _check(self.lib.phet_run_pipeline(self.h, ptr, 1 if is_device else 0, N, G, dtype, C.byref(pp)))
        self.N, self.G = N, G
        return idx_out

    def fy_apply(self, draws: np.ndarray, n: int, take: int | None = None) -> np.ndarray:
        """Permutations from swap targets on the device: draws (count, >= n - 1) uint16 / uint32 -> uint32 (count, take)."""
        draws = np.ascontiguousarray(draws)
        take = int(n) if take is None else int(take)
        out = np.empty((draws.shape[0], take), dtype=np.uint32)
        _check(self.lib.phet_fy_apply(self.h, _ptr(draws), draws.dtype.itemsize, draws.shape[1], int(n), draws.shape[0], take, _ptr(out)))
        return out

    def draw_index_sets(self, mt: "MtStream", S: int, m: int) -> np.ndarray:
        """Index sets drawn from ``mt`` as the reference draws them (phet.py:397-402, replace=False), installed on
        the device; returns them as int32 (S, 2, m) original row numbers."""
        idx = np.empty((int(S), 2, int(m)), dtype=np.int32)
        _check(self.lib.phet_draw_index_sets(self.h, mt.h, int(S), int(m), _ptr(idx)))
        self.S, self.m = int(S), int(m)
        return idx

    def phase_ms(self):
        """(normalise, Step 1, Step 3) device milliseconds of the last device-matrix pipeline call (-1: not available)."""
        out = (C.c_float * 3)()
        _check(self.lib.phet_phase_ms(self.h, out))
        return float(out[0]), float(out[1]), float(out[2])

    def replay_stats(self):
        out = (C.c_double * 2)()
        _check(self.lib.phet_replay_stats(self.h, out))
        return dict(draw_thread_s=float(out[0]), issue_wait_s=float(out[1]))

    def fisher(self, S_total):
        _check(self.lib.phet_fisher(self.h, S_total))

    def o_minmax(self):
        out = (C.c_double * 2)()
        _check(self.lib.phet_o_minmax(self.h, out))
        return float(out[0]), float(out[1])

    def bin(self, edges: np.ndarray):
        e = np.ascontiguousarray(edges, dtype=np.float64)
        B = e.size - 1
        counts = np.zeros(B, dtype=np.int64)
        _check(self.lib.phet_bin(self.h, _ptr(e), B, _ptr(counts)))
        return counts

    def combine(self, weights: np.ndarray, fetch=True):
        w = np.ascontiguousarray(weights, dtype=np.float64)
        out = np.empty(self.G, dtype=np.float64) if fetch else None
        _check(self.lib.phet_combine(self.h, _ptr(w), w.size, _ptr(out) if fetch else None))
        return out

    def finalize(self, S_total, weights: np.ndarray, fetch=True):
        """Fisher + uniform binning of o + combine on the device in one call: (scores, bin counts, edges)."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        B = w.size
        scores = np.empty(self.G, dtype=np.float64) if fetch else None
        counts = np.zeros(B, dtype=np.int64)
        edges = np.empty(B + 1, dtype=np.float64)
        _check(self.lib.phet_finalize(self.h, int(S_total), _ptr(w), B, _ptr(scores) if fetch else None, _ptr(counts),
                                      _ptr(edges)))
        return scores, counts, edges

    def finalize_async(self, S_total, weights: np.ndarray):
        """The work of finalize() queued on the stream without a host synchronisation; results stay on the device
        until fetch_results()."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        self._n_bins = w.size
        _check(self.lib.phet_finalize_async(self.h, int(S_total), _ptr(w), w.size))

    def fetch_results(self, scores=True):
        """(scores or None, bin counts, edges) of the last finalized fit; synchronises."""
        B = self._n_bins
        sc = np.empty(self.G, dtype=np.float64) if scores else None
        counts = np.zeros(B, dtype=np.int64)
        edges = np.empty(B + 1, dtype=np.float64)
        _check(self.lib.phet_fetch_results(self.h, _ptr(sc) if scores else None, _ptr(counts), _ptr(edges)))
        return sc, counts, edges

    def device_ptr(self, which):
        p = C.c_void_p()
        n = C.c_int64()
        _check(self.lib.phet_device_ptr(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def read(self, which, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        _check(self.lib.phet_read_buffer(self.h, which, _ptr(out), out.nbytes))
        return out

    def write(self, which, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        _check(self.lib.phet_write_buffer(self.h, which, _ptr(arr), arr.nbytes))

    def step3_path(self):
        return {0: "none", 1: "full-warp sort", 2: "half-warp sort", 3: "bound-and-prune", 4: "scatter bound-and-prune"}[int(self.lib.phet_step3_path(self.h))]

    def launch_count(self):
        return int(self.lib.phet_launch_count(self.h))

    def step1_geometry(self):
        out = (C.c_int32 * 5)()
        _check(self.lib.phet_step1_geometry(self.h, out))
        path = "pair-per-thread" if out[4] & 4 else ("half-warp" if out[4] & 2 else "full-warp")
        return dict(grid=out[0], block=out[1], smem=out[2], E=out[3], staged=bool(out[4] & 1), path=path)
