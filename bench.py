"""Benchmark of the PHet fit_predict hot path on B200 (driver contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--permutations reference|device]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the metric is quoted on; fits one GPU): synthetic
50,000 cells x 30,000 genes float32, 25k/25k classes, S = 1000 subsamples, m = 158, iqr-range
25-75, feature-weight [0.4, 0.3, 0.2, 0.1].  A "step" is one complete fit: z-score + transpose,
Step 1 (all gene x subsample units), Step 3 (KS over the permuted class columns), Fisher, binning, combine.
Metric: gene x subsample statistics per second = G*S / time per fit.

Random draws (--permutations, default "reference"): the index sets and the 2 G Step-3 permutations are the ones the
reference would draw from np.random seeded with 12345 (phet.py:401-402, 486-487), replayed natively.

  value   inputs resident in HBM when the timed region starts: the raw matrix, the index sets and (reference mode)
          the permutation tables -- all "host-generated index sets" in north-star terms.  The K fits are issued back to
          back (phet_finalize_async: every fit is complete on the device, scores in PHET_BUF_SCORES; the last one is
          copied out before the closing event), timed with CUDA events on the library's stream, max over ranks
  e2e     the user's call: PHet(...).fit_predict(X_host, y) from pinned host memory, np.random seeded just before;
          the draws (host scan of the MT19937 stream, overlapped with the device) are INSIDE the timed region
  e2e_device_perm / value_device_perm: the same with permutation="device" (a keyed bijection on the GPU instead of
          the reference's stream: NOT the reference's scores) -- round 1's headline, kept for continuity
  N > 1   genes sharded over ranks (every stage shards; --shard subsamples gives the north-star layout
          instead), one NCCL all-reduce; total work fixed -> "scaling": "strong"; rank 0 checks that the N-rank
          scores equal a one-rank fit bit for bit

The working set (6 GB raw + 6 GB gene-major) is far larger than the 126 MB L2, so no explicit
L2 flush is needed between timed steps (config.l2 says so).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg4": dict(name="synthetic 50k cells x 30k genes, 1000 subsamples (BASELINE configs[3])", N=50000, G=30000,
                 S=1000),
    "cfg3": dict(name="synthetic 10k cells x 30k genes, 1000 subsamples (BASELINE configs[2])", N=10000, G=30000,
                 S=1000),
    # BASELINE configs[4] (HBM-capacity sizing): the whole matrix (48 GB raw + 48 GB gene-major) fits one B200;
    # "cfg5r8" is the share of ONE rank when the genes are sharded over the 8 GPUs of a box (7,500 genes)
    "cfg5": dict(name="synthetic 200k cells x 60k genes, 5000 subsamples (BASELINE configs[4])", N=200000, G=60000,
                 S=5000),
    "cfg5r8": dict(name="one rank's gene slice (7,500 of 60k genes) of BASELINE configs[4]: 200k cells, 5000 subsamples",
                   N=200000, G=7500, S=5000),
    "tiny": dict(name="synthetic 2k cells x 512 genes, 64 subsamples (debug)", N=2000, G=512, S=64),
}
FEATURE_WEIGHT = [0.4, 0.3, 0.2, 0.1]
PERCENTILES = (25, 75)
SEED = 12345


def ncu_step1_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the staged k_step1p<float> launch at config 4, read from the
    newest tracked `ncu --set full` summary under profiles/ (profiles/summarize_ncu.py format).  (bytes, file) or
    (None, None)."""
    import glob
    import re
    best = (None, None)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_v*.txt")),
                       key=lambda f: [int(x) for x in re.findall(r"\d+", os.path.basename(f))]):
        rd = wr = None
        inside = False
        for line in open(path):
            if line.startswith("kernel:"):
                if inside and rd is not None and wr is not None:
                    break
                inside = "k_step1p<float" in line
                rd = wr = None
                continue
            if not inside:
                continue
            m_ = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m_:
                v = float(m_.group(2)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(m_.group(3), 1.0)
                if m_.group(1) == "read":
                    rd = v
                else:
                    wr = v
        if rd is not None and wr is not None and 5.5e9 < rd + wr < 7e9:   # a config-4 launch (6.0 GB algorithmic)
            best = (int(rd + wr), os.path.relpath(path, ROOT))
    return best


# ----------------------------------------------------------------------------------------------
def synth_matrix_device(N, G, device, seed=0):
    """SURVEY 8(d) synthetic data, generated on the device: gamma(2,1) float32; first N//2 rows
    control; first 5 % of genes mean-shifted (+1) on case rows, next 5 % dispersion-shifted (x2
    about the gene mean).  Identical on every rank (fixed-seed CUDA generator)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    X = torch.empty((N, G), dtype=torch.float32, device=device)
    rows = max(1, (256 << 20) // (4 * G))
    for r0 in range(0, N, rows):  # gamma(2,1) = -log(u1) - log(u2); chunked to bound temporaries
        r1 = min(N, r0 + rows)
        u = torch.rand((2, r1 - r0, G), generator=gen, device=device, dtype=torch.float32).clamp_(min=1e-12)
        X[r0:r1] = -(torch.log(u[0]) + torch.log(u[1]))
        del u
    k = max(G // 20, 1)
    n_ctrl = N // 2
    X[n_ctrl:, :k] += 1.0
    mu = X[:, k:2 * k].mean(0)
    X[n_ctrl:, k:2 * k] = (X[n_ctrl:, k:2 * k] - mu) * 2 + mu
    return X


def draw_index_sets(N, S, m):
    """Host-generated index sets exactly as the reference draws them (phet.py:401-402) after np.random.seed(SEED):
    the library's native replay of the stream (bit-identical to np.random.choice, tests/test_mt_replay_cpu.py).
    Returns the rows, the index sets and the stream positioned where the reference's Step-3 draws begin."""
    from phet_b200 import _capi
    n_ctrl = N // 2
    ctrl = np.arange(n_ctrl)
    case = np.arange(n_ctrl, N)
    np.random.seed(SEED)
    mt = _capi.MtStream.from_numpy()
    idx = np.empty((S, 2, m), dtype=np.int32)
    for s in range(S):
        idx[s, 0] = ctrl[mt.permutation(len(ctrl), take=m)]
        idx[s, 1] = case[mt.permutation(len(case), take=m)]
    return ctrl.astype(np.int32), case.astype(np.int32), idx, mt


class ClockSampler:
    """nvidia-smi sampler for the timed region (recipe in B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
class Fit:
    """One configured fit driven through the C ABI (phet_b200._capi.Context)."""

    def __init__(self, ctx, wl, ctrl, case, idx, rank, world, dist, shard="genes", perms="reference"):
        from phet_b200 import _capi, hostmath
        from phet_b200.distributed import shard_range
        self.capi, self.ctx, self.wl, self.rank, self.world, self.dist = _capi, ctx, wl, rank, world, dist
        self.N, self.G, self.S = wl["N"], wl["G"], wl["S"]
        self.ctrl, self.case, self.idx = ctrl, case, idx
        self.m = idx.shape[2]
        self.min_c = min(len(ctrl), len(case))
        self.n_slices, self.n_last = hostmath.step3_slices(self.min_c, self.m)
        self.tab_full = hostmath.ks_exact_table(self.m)
        self.tab_last = hostmath.ks_exact_table(self.n_last)
        (j0, k0, g0), (j1, k1, g1) = hostmath.quantile_positions(self.m, PERCENTILES, dtype=np.float32)
        prm = _capi.Step1Params()
        prm.test = _capi.TEST_T if self.m < 30 else _capi.TEST_Z
        prm.ln_beta = hostmath.ln_beta_half(2.0 * self.m - 2.0) if prm.test == _capi.TEST_T else 0.0
        prm.p_flush_below = 0.0
        prm.q = _capi.QuantilePos(j0, k0, j1, k1, g0, g1)
        prm.capture_debug = 0
        self.prm = prm
        self.shard = shard
        self.perms = perms
        if shard == "subsamples":   # north star: whole matrix per rank, slice of the index sets
            self.gl_lo, self.gl_hi = 0, self.G
            self.s_lo, self.s_hi = shard_range(self.S, rank, world)
        else:                       # genes: every rank owns a gene slice, all subsamples
            self.gl_lo, self.gl_hi = shard_range(self.G, rank, world)
            self.s_lo, self.s_hi = 0, self.S
        self.g_lo, self.g_hi = shard_range(self.G, rank, world)
        self.weights = np.array(FEATURE_WEIGHT) / np.sum(FEATURE_WEIGHT)
        self.hostmath = hostmath

    def _pipeline(self, X_dev_ptr, **kw):
        c, capi = self.ctx, self.capi
        return c.run_pipeline(X_dev_ptr, self.N, self.G, capi.PHET_F32, self.prm, self.m, self.tab_full, self.tab_last,
                              self.n_last, g_range=(self.gl_lo, self.gl_hi), s_range=(self.s_lo, self.s_hi),
                              g3_range=(self.g_lo, self.g_hi), is_device=True, **kw)

    def setup_resident(self, X_dev_ptr, mt):
        """Everything `value` treats as resident input: class layout, index-set tables and -- reference mode -- this
        rank's permutation tables, drawn once from the stream (where the reference's Step-3 draws begin) and kept."""
        self.ctx.set_classes(self.ctrl, self.case)
        self.ctx.set_index_sets(self.idx)
        if self.perms == "reference":
            # (N > 1: a rank needs the stream up to its own genes only; the final state of `mt` is not used here)
            self._pipeline(X_dev_ptr, mt=mt, replay_perms=True, keep_perm_tables=True, stop_after_own_genes=self.world > 1)
            self.ctx.sync()

    def finish(self, fetch=True):
        """All-reduce (N > 1) + Fisher / binning / combine, all queued on the context's stream; fetch=True also copies
        the scores and bin counts out (one host synchronisation)."""
        c = self.ctx
        if self.dist is not None:
            from phet_b200.distributed import allreduce_device_buffer
            allreduce_device_buffer(c, self.capi.BUF_ACC_O, self.dist)   # sum -2 ln p | sum delta | o in one all-reduce
        c.finalize_async(self.S, self.weights)
        if not fetch:
            return None
        scores, self.counts, _ = c.fetch_results()
        return scores

    def step_resident(self, X_dev_ptr, perms=None, fetch=True):
        """Timed unit for `value`: raw matrix + index sets (+ permutation tables) already in HBM.  fetch=False leaves the
        scores on the device and does not synchronise, so that consecutive fits run back to back on the GPU."""
        if (perms or self.perms) == "reference":
            self._pipeline(X_dev_ptr, resident_perms=True)
        else:
            self._pipeline(X_dev_ptr, seed=SEED)
        return self.finish(fetch)


# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(wl, X_cols, ctrl, case, idx, g3=24, s_sub=None, threads=1):
    """Time the oracle port (numpy/scipy restatement of the reference, single-threaded like the
    reference) on a bounded sample of the workload and extrapolate linearly in G and S (both
    loops of phet.py are exactly linear in them).  X_cols: (N, g_sub) float32 host columns."""
    from oracle import phet_oracle as po
    N, G, S = wl["N"], wl["G"], wl["S"]
    g_sub = X_cols.shape[1]
    m = idx.shape[2]
    s_sub = s_sub or min(S, 40)
    t0 = time.perf_counter()
    Xz = po.zscore_nan_to_num(X_cols)
    t_z = time.perf_counter() - t0
    t0 = time.perf_counter()
    for s in range(s_sub):
        po.step1_subsample(Xz, idx[s, 0], idx[s, 1], PERCENTILES, m, library=True)
    t_s1 = time.perf_counter() - t0
    rng = np.random.RandomState(1)
    perm_i = np.stack([rng.permutation(len(ctrl)) for _ in range(g3)])
    perm_j = np.stack([rng.permutation(len(case)) for _ in range(g3)])
    t0 = time.perf_counter()
    po.step3_o_library(Xz[:, :g3], ctrl, case, perm_i, perm_j, m)
    t_s3 = time.perf_counter() - t0
    t_fit = t_z * (G / g_sub) + t_s1 * (G / g_sub) * (S / s_sub) + t_s3 * (G / g3)
    sample = ("oracle port of phet.py on %d of %d genes x all %d cells: zscore %.2fs; Step 1 %d of %d subsamples "
              "%.2fs; Step 3 (scipy ks_2samp loop) %d genes %.2fs; extrapolated linearly in G and S to a full fit "
              "of %.0f s" % (g_sub, G, N, t_z, s_sub, S, t_s1, g3, t_s3, t_fit))
    return {"value": G * S / t_fit, "unit": "gene*subsample/s", "cores": threads, "kind": "port", "sample": sample,
            "fit_seconds_extrapolated": t_fit, "step1_units_per_s": g_sub * s_sub / t_s1}


# ----------------------------------------------------------------------------------------------
def cpu_reference_class_sample(wl, X_cols, y, ctrl, case, idx):
    """Time the UNMODIFIED reference class (oracle/_ref/model/phet.py under oracle/ref_shim.py; vendored by
    oracle/make_ref.sh) on bounded sub-problems of the workload -- all N cells, (genes, subsamples) in (16, 16),
    (16, 48), (64, 16) -- and extrapolate with T = d + a g + b s: a is Step 3 + z-score per gene (the bulk of a fit:
    159 ks_2samp calls per gene), b the per-subsample overhead (two np.random.choice over 25k rows).  The term
    proportional to genes x subsamples (the column-wise iqr / ztest work of Step 1) cannot be resolved from the
    monolithic class within the time bound -- at 64 genes it is ~1 % of a run -- so it is taken from the oracle
    port's Step 1 (the same scipy calls) on 1,024 genes.  Single-threaded like the reference.  None when the vendored
    copy is absent."""
    from oracle import phet_oracle as po
    from oracle import ref_shim
    if not ref_shim.reference_available():
        return None
    PHet = ref_shim.load_reference_phet()
    N, G, S = wl["N"], wl["G"], wl["S"]
    pts = [(16, 16), (16, 48), (64, 16)]
    T = []
    for g, s_ in pts:
        np.random.seed(SEED)
        t0 = time.perf_counter()
        # (the class's own keyword is `percentiles`, phet.py:33-40; train.py:474 passes iqr_range= and fails on it)
        PHet(normalize="zscore", percentiles=PERCENTILES, num_subsamples=s_, subsampling_size="sqrt",
             feature_weight=FEATURE_WEIGHT).fit_predict(np.array(X_cols[:, :g], copy=True), y)
        T.append(time.perf_counter() - t0)
    A = np.array([[1.0, g, s_] for g, s_ in pts])
    d, a, b = np.linalg.solve(A, np.array(T))
    g_sub = X_cols.shape[1]
    s_sub = min(idx.shape[0], 24)
    Xz = po.zscore_nan_to_num(X_cols)
    t0 = time.perf_counter()
    for s_ in range(s_sub):
        po.step1_subsample(Xz, idx[s_, 0], idx[s_, 1], PERCENTILES, idx.shape[2], library=True)
    c = (time.perf_counter() - t0) / (g_sub * s_sub)
    t_fit = float(d + a * G + b * S + c * G * S)
    sample = ("unmodified reference PHet.fit_predict (oracle/_ref) on all %d cells x (genes, subsamples) in %s: %s s -> "
              "T = d + a g + b s with d=%.3g a=%.3g b=%.3g; gene x subsample term from the oracle port's Step 1 on %d genes x "
              "%d subsamples: c=%.3g s/unit; full fit (%d genes, %d subsamples) = %.0f s (Step 3 + zscore %.0f, draws %.0f, "
              "Step 1 %.0f)" % (N, pts, ["%.2f" % t for t in T], d, a, b, g_sub, s_sub, c, G, S, t_fit, a * G, b * S, c * G * S))
    return {"value": G * S / t_fit, "unit": "gene*subsample/s", "cores": 1, "kind": "reference", "sample": sample,
            "fit_seconds_extrapolated": t_fit}


def shared_pinned_matrix(X_dev, N, G, rank, world, local_rank, dist):
    """ONE host copy of the matrix per node, pinned by every rank (cudaHostRegister on a /dev/shm mapping), instead
    of one pinned 4 N G-byte copy per rank."""
    import torch
    path = "/dev/shm/phet_bench_X_%s_%dx%d.f32" % (os.environ.get("MASTER_PORT", "solo"), N, G)
    if rank == 0:
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(N, G))
        rows = max(1, (512 << 20) // (4 * G))
        for r0 in range(0, N, rows):
            mm[r0:r0 + rows] = X_dev[r0:r0 + rows].cpu().numpy()
        mm.flush()
        del mm
    if dist is not None:
        dist.barrier()
    Xh = np.load(path, mmap_mode="r+")
    rc = torch.cuda.cudart().cudaHostRegister(Xh.ctypes.data, Xh.nbytes, 0)
    if int(rc) != 0:
        raise RuntimeError("cudaHostRegister failed: %s" % rc)
    if dist is not None:
        dist.barrier()
    if rank == 0:
        os.unlink(path)   # the mappings keep it alive
    return Xh


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="genes", choices=["subsamples", "genes"],
                    help="N > 1: genes (default) = every rank owns a gene slice: normalise, Step 1 and Step 3 all "
                         "shard and nothing is replicated; subsamples = north-star layout (matrix replicated, index "
                         "sets sliced, Step 3 gene-sliced).  Both end in the same NCCL all-reduce.")
    ap.add_argument("--permutations", default=None, choices=["reference", "device"],
                    help="Step-3 permutations of the headline numbers (value, e2e): reference = the reference's np.random "
                         "stream, replayed natively (default; config 5 defaults to device: its 1.2e10 draws take seconds "
                         "on the host); the other mode is reported beside it (value_device_perm / e2e_device_perm)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-at-n", action="store_true", help="N > 1: rank 0 still times the CPU baseline (default: N = 1 only)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    N, G, S = wl["N"], wl["G"], wl["S"]
    perms = args.permutations or ("device" if args.workload.startswith("cfg5") else "reference")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from phet_b200 import hostmath
    m = hostmath.subsample_size(N // 2, N - N // 2, "sqrt")
    config = {"workload": wl["name"], "cells": N, "genes": G, "subsamples": S, "m": m, "iqr_range": list(PERCENTILES),
              "feature_weight": FEATURE_WEIGHT, "storage": "f32",
              "step3_permutations": ("reference: np.random stream (seed 12345) replayed natively; value = tables resident in HBM, "
                                     "e2e = drawn inside the timed region" if perms == "reference" else
                                     "device: keyed bijection evaluated on the GPU (not the reference's stream)"),
              "l2": "working set %.0f GB >> 126 MB L2, no flush needed" % (2 * 4.0 * N * G / 1e9),
              "parallelism": "%s-shard x%d" % (args.shard, world)}

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference_arm(args, wl, config, m)
        return

    import torch
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=dev)
        dist = tdist
    from phet_b200 import PHet, _capi
    # a dedicated (non-default) torch stream: the C library launches on it, torch events and NCCL
    # collectives are enqueued on it, so device-side timing covers exactly the library's kernels
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = _capi.Context(local_rank, stream.cuda_stream)

    X = synth_matrix_device(N, G, dev, seed=0)
    ctrl, case, idx, mt = draw_index_sets(N, S, m)
    fit = Fit(ctx, wl, ctrl, case, idx, rank, world, dist, shard=args.shard, perms=perms)
    t0 = time.perf_counter()
    fit.setup_resident(X.data_ptr(), mt)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor(np.atleast_1d(np.asarray(v, dtype=np.float64)), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = t.cpu().numpy()
        return float(out[0]) if np.ndim(v) == 0 else out

    def timed_resident(mode, steps, warmup):
        for _ in range(warmup):
            sc = fit.step_resident(X.data_ptr(), mode)
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = []
        barrier()
        tw = time.perf_counter()
        e0.record()
        for k in range(steps):
            # every fit is complete on the device (scores in PHET_BUF_SCORES); only the last one is copied out, so the
            # host issues fit k + 1 while fit k runs and the GPU stays busy back to back (no per-fit synchronisation)
            sc = fit.step_resident(X.data_ptr(), mode, fetch=(k == steps - 1))
        e1.record()
        barrier()
        tw = time.perf_counter() - tw
        phases.append(ctx.phase_ms())   # in-stream events of the last fit
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        ph = max_over_ranks(np.array(phases).mean(0))
        return sc, ms, ph, ctx.launch_count() - l0, tw / steps

    sampler = ClockSampler(local_rank)
    sampler.start()
    scores, ms_per_step, phase, launches, wall_s = timed_resident(perms, args.steps, args.warmup)
    clocks = sampler.stop()
    counts_main = [int(c) for c in fit.counts]
    value = G * S / (ms_per_step * 1e-3)
    other = "device" if perms == "reference" else "reference"
    value_other = None
    if not (other == "reference" and args.workload.startswith("cfg5")):
        if other == "reference":
            fit.perms = "reference"
            fit.setup_resident(X.data_ptr(), draw_index_sets(N, S, m)[3])
        _, ms_o, ph_o, _, _ = timed_resident(other, max(1, min(args.steps, 3)), 2)
        value_other = {"value": G * S / (ms_o * 1e-3), "ms_per_step": ms_o, "bin_counts": [int(c) for c in fit.counts],
                       "phases_ms": {"normalize_transpose": ph_o[0], "step1": ph_o[1], "step3": ph_o[2]}}

    # ---- e2e: the user's call, PHet(...).fit_predict(X_host, y), host matrix pinned once per node
    e2e = e2e_other = None
    scores_e2e = None
    if not args.no_e2e:
        X_host = shared_pinned_matrix(X, N, G, rank, world, local_rank, dist)
        y = np.zeros(N, dtype=np.int64)
        y[N // 2:] = 1

        def class_fit(mode, distributed):
            np.random.seed(SEED)
            est = PHet(normalize="zscore", iqr_range=PERCENTILES, num_subsamples=S, feature_weight=FEATURE_WEIGHT,
                       permutation=mode, device=local_rank, distributed=distributed, shard=args.shard)
            return est, est.fit_predict(X_host, y)

        def timed_e2e(mode):
            steps = max(1, min(args.steps, 3))
            est, sc = class_fit(mode, "auto")
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                est, sc = class_fit(mode, "auto")
            barrier()
            dt = max_over_ranks((time.perf_counter() - t0) / steps)
            g_loc = fit.gl_hi - fit.gl_lo
            draws = 0
            if mode == "reference":   # swap targets uploaded: 2 bytes each (classes <= 65536 cells), S index rounds + own genes
                eb = 2 if max(len(ctrl), len(case)) <= 65536 else 4
                draws = (S + (fit.g_hi - fit.g_lo)) * (len(ctrl) + len(case)) * eb
            return est, sc, {"value": G * S / dt, "unit": "gene*subsample/s", "ms_per_step": dt * 1e3,
                             "h2d_bytes_per_step": int(world * (4 * N * g_loc + draws + 12 * N)),
                             "d2h_bytes_per_step": int(world * (8 * G + 8 * G + idx.nbytes + 64)),
                             "h2d_gbs_aggregate": world * (4 * N * g_loc + draws) / dt / 1e9,
                             "timings_rank0": {k: float(v) for k, v in est.timings_.items()}}

        est, scores_e2e, e2e = timed_e2e(perms)
        e2e["what"] = ("np.random.seed(12345); PHet(...).fit_predict(X_host, y) through the drop-in class: class layout, "
                       "%s, H2D of the pinned matrix in gene blocks overlapped with normalise / Step 1 / Step 3, "
                       "all-reduce, finalize, scores D2H; wall clock, max over ranks"
                       % ("index sets and 2 G permutations drawn from the np.random stream on host threads (scan) + device "
                          "(swaps), overlapped with the device work" if perms == "reference" else
                          "index sets drawn from the np.random stream, Step-3 permutations by the device bijection"))
        assert np.allclose(scores_e2e[:, 0], scores, rtol=1e-9, atol=0), "class path and C-ABI path disagree"
        if world > 1 and rank == 0:
            # gene sharding is exact: the N-rank fit must equal a one-rank fit of the same data bit for bit
            _, single = class_fit(perms, False)
            assert np.array_equal(single, scores_e2e), "N-rank scores differ from the one-rank fit"
            e2e["equals_one_rank_fit_bitwise"] = True
        barrier()
        if value_other is not None:
            _, _, e2e_other = timed_e2e(other)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_step1), live CUDA-event duration on its stream
    geom = ctx.step1_geometry()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    s_local = fit.s_hi - fit.s_lo
    g_local = (fit.gl_hi - fit.gl_lo)
    n_pad = (N + 3) // 4 * 4
    hbm_bytes = g_local * n_pad * 4 + 16 * g_local + s_local * 2 * m * 4   # gene rows staged once + accumulators + idx
    smem_bytes = g_local * s_local * 2 * m * 4                             # gathered values
    step1_s = phase[1] * 1e-3
    kname = {"pair-per-thread": "k_step1p<float>", "half-warp": "k_step1h<float,E=%d>" % geom["E"],
             "full-warp": "k_step1<float,E=%d>" % geom["E"]}[geom["path"]]
    traffic, traffic_src = ncu_step1_traffic() if (args.workload == "cfg4" and world == 1 and geom["path"] == "pair-per-thread" and geom["staged"]) else (None, None)
    roofline = {"bound": "hbm", "kernel": "%s (staged=%s)" % (kname, geom["staged"]),
                "achieved": hbm_bytes / step1_s / 1e9, "peak": peak, "unit": "GB/s",
                "frac": hbm_bytes / step1_s / 1e9 / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the newest tracked `ncu --set full` summary
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "launch_ms": phase[1], "algorithmic_hbm_bytes_per_launch": int(hbm_bytes),
                "smem_gather_gbs": smem_bytes / step1_s / 1e9,
                "smem_gather_peak_gbs": 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9,
                "note": "the Step-1 kernel is not HBM bound: each class vector is staged once per gene (traffic == "
                        "algorithmic bytes) and then serves ~1000 samples from SMEM; its limiter is instruction issue "
                        "(see the ncu summary named in traffic_source)",
                "grid": geom["grid"], "block": geom["block"], "smem_per_cta": geom["smem"]}
    cpu = None
    # the CPU baseline is rank 0's job, at N = 1 (or at N > 1 when asked for: --cpu-at-n, e.g. config 5 on 8 GPUs)
    if not args.no_cpu and rank == 0 and (world == 1 or args.cpu_at_n):
        g_sub = min(G, 1024)
        cols = X[:, :g_sub].contiguous().cpu().numpy()
        y = np.zeros(N, dtype=np.int64)
        y[N // 2:] = 1
        cpu = cpu_reference_class_sample(wl, cols, y, ctrl, case, idx) or cpu_reference_sample(wl, cols, ctrl, case, idx)
    line = {"metric": "PHet fit gene*subsample statistics per second", "value": value, "unit": "gene*subsample/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 statistics over f32-staged z-scores",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "phases_ms": {"normalize_transpose": phase[0], "step1": phase[1], "step3": phase[2],
                          "allreduce_finalize": ms_per_step - float(np.sum(phase))},
            "wall_ms_per_step": wall_s * 1e3, "bin_counts": counts_main, "score_sum": float(np.sum(scores)),
            "resident_setup_s": setup_s,
            ("value_device_perm" if other == "device" else "value_reference_perm"): value_other,
            ("e2e_device_perm" if other == "device" else "e2e_reference_perm"): e2e_other}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_reference_arm(args, wl, config, m):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores -- the unmodified
    class vendored under oracle/_ref when present (kind "reference"), else the oracle port -- single-threaded like
    the reference (PHet never uses --num-jobs), each step a bounded sample of the same workload."""
    N, G, S = wl["N"], wl["G"], wl["S"]
    g_sub = min(G, 1024)
    rng = np.random.default_rng(0)
    X = rng.gamma(2.0, 1.0, (N, g_sub)).astype(np.float32)
    k = max(X.shape[1] // 20, 1)
    X[N // 2:, :k] += 1.0
    mu = X[:, k:2 * k].mean(0)
    X[N // 2:, k:2 * k] = (X[N // 2:, k:2 * k] - mu) * 2 + mu
    y = np.zeros(N, dtype=np.int64)
    y[N // 2:] = 1
    ctrl, case, idx, _ = draw_index_sets(N, min(S, 64), m)
    res = None
    times = []
    for i in range(args.warmup + args.steps):
        res = cpu_reference_class_sample(wl, X, y, ctrl, case, idx) or cpu_reference_sample(wl, X, ctrl, case, idx, g3=8, s_sub=min(S, 24))
        if i >= args.warmup:
            times.append(res["fit_seconds_extrapolated"])
        if res["kind"] == "reference" and len(times) >= 1 and i >= 1:
            break   # one bounded sample of the unmodified class is ~20-60 s of CPU: two passes are enough
    t_fit = float(np.mean(times)) if times else res["fit_seconds_extrapolated"]
    value = G * S / t_fit
    res["value"] = value
    line = {"impl": "reference", "metric": "PHet fit gene*subsample statistics per second", "value": value,
            "unit": "gene*subsample/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_fit * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 input, numpy/scipy arithmetic", "data": "synthetic", "config": config,
            "cpu_baseline": res, "gpu_launches": 0,
            "e2e": {"value": value, "unit": "gene*subsample/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
