"""Benchmark of the PHet fit_predict hot path on B200 (driver contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the metric is quoted on; fits one GPU): synthetic
50,000 cells x 30,000 genes float32, 25k/25k classes, S = 1000 subsamples, m = 158, iqr-range
25-75, feature-weight [0.4, 0.3, 0.2, 0.1].  A "step" is one complete fit: z-score + transpose,
Step 1 (all gene x subsample units), Step 3 (KS, device permutations), Fisher, binning, combine.
Metric: gene x subsample statistics per second = G*S / time per fit.

  value   raw matrix and index sets already resident in HBM when the timed region starts
  e2e     same fit through the C ABI with HOST buffers (pinned matrix + index sets H2D, scores D2H)
  N > 1   genes sharded over ranks (every stage shards; --shard subsamples gives the north-star layout
          instead), two NCCL all-reduces; total work fixed -> "scaling": "strong"

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
NCU_STEP1_DRAM_BYTES = 6003228000 + 7337000   # k_step1p<float,16,1>, config 4, one launch (profiles/r1_ncu_full_v6.txt)


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
    """Host-generated index sets exactly as the reference draws them (phet.py:401-402)."""
    n_ctrl = N // 2
    ctrl = np.arange(n_ctrl)
    case = np.arange(n_ctrl, N)
    np.random.seed(SEED)
    idx = np.empty((S, 2, m), dtype=np.int32)
    for s in range(S):
        idx[s, 0] = np.random.choice(a=ctrl, size=m, replace=False)
        idx[s, 1] = np.random.choice(a=case, size=m, replace=False)
    return ctrl.astype(np.int32), case.astype(np.int32), idx


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

    def __init__(self, ctx, wl, ctrl, case, idx, rank, world, dist, shard="subsamples"):
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
        if shard == "subsamples":   # north star: whole matrix per rank, slice of the index sets
            self.gl_lo, self.gl_hi = 0, self.G
            self.s_lo, self.s_hi = shard_range(self.S, rank, world)
        else:                       # genes: every rank owns a gene slice, all subsamples
            self.gl_lo, self.gl_hi = shard_range(self.G, rank, world)
            self.s_lo, self.s_hi = 0, self.S
        self.g_lo, self.g_hi = shard_range(self.G, rank, world)
        self.weights = np.array(FEATURE_WEIGHT) / np.sum(FEATURE_WEIGHT)
        self.hostmath = hostmath

    def setup_resident(self):
        self.ctx.set_classes(self.ctrl, self.case)

    def finish(self, fetch=True):
        c = self.ctx
        if self.dist is not None:
            from phet_b200.distributed import allreduce_device_buffer
            allreduce_device_buffer(c, self.capi.BUF_ACC_O, self.dist)   # sum -2 ln p | sum delta | o in one all-reduce
        scores, self.counts, _ = c.finalize(self.S, self.weights, fetch=fetch)
        return scores

    def step_resident(self, X_dev_ptr, events=None):
        """Timed unit for `value`: raw matrix + index sets already in HBM."""
        c, capi = self.ctx, self.capi
        rec = (lambda k: events[k].record()) if events else (lambda k: None)
        rec(0)
        if self.shard == "subsamples" or self.world == 1:   # one rank: both layouts are the same work
            c.load_matrix_device(X_dev_ptr, self.N, self.G, capi.PHET_F32, capi.NORM_ZSCORE, capi.PHET_F32)
            rec(1)
            c.step1(self.s_lo, self.s_hi, self.prm)
            rec(2)
            c.step3(self.g_lo, self.g_hi, self.m, self.tab_full, self.tab_last, self.n_last, seed=SEED)
            rec(3)
        else:
            c.run_pipeline(X_dev_ptr, self.N, self.G, capi.PHET_F32, self.prm, self.m, self.tab_full, self.tab_last,
                           self.n_last, g_range=(self.gl_lo, self.gl_hi), s_range=(self.s_lo, self.s_hi),
                           g3_range=(self.g_lo, self.g_hi), seed=SEED, is_device=True)
            rec(1), rec(2), rec(3)
        out = self.finish()
        rec(4)
        return out

    def step_e2e(self, X_host):
        """Timed unit for `e2e`: host matrix (pinned) and host index sets in, host scores out, through
        the gene-block pipeline (H2D of block b+1 overlaps the kernels of block b)."""
        c, capi = self.ctx, self.capi
        c.set_classes(self.ctrl, self.case)
        c.set_index_sets(self.idx)
        c.run_pipeline(X_host, self.N, self.G, capi.PHET_F32, self.prm, self.m, self.tab_full, self.tab_last,
                       self.n_last, g_range=(self.gl_lo, self.gl_hi), s_range=(self.s_lo, self.s_hi),
                       g3_range=(self.g_lo, self.g_hi), seed=SEED)
        return self.finish()


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
                         "sets sliced, Step 3 gene-sliced).  Both end in the same two NCCL all-reduces.")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    N, G, S = wl["N"], wl["G"], wl["S"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from phet_b200 import hostmath
    m = hostmath.subsample_size(N // 2, N - N // 2, "sqrt")
    config = {"workload": wl["name"], "cells": N, "genes": G, "subsamples": S, "m": m, "iqr_range": list(PERCENTILES),
              "feature_weight": FEATURE_WEIGHT, "storage": "f32", "step3_permutations": "device",
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
    from phet_b200 import _capi
    # a dedicated (non-default) torch stream: the C library launches on it, torch events and NCCL
    # collectives are enqueued on it, so device-side timing covers exactly the library's kernels
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = _capi.Context(local_rank, stream.cuda_stream)

    X = synth_matrix_device(N, G, dev, seed=0)
    ctrl, case, idx = draw_index_sets(N, S, m)
    fit = Fit(ctx, wl, ctrl, case, idx, rank, world, dist, shard=args.shard)
    fit.setup_resident()
    ctx.set_index_sets(idx)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        scores = fit.step_resident(X.data_ptr())
    barrier()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for k in range(args.steps):
        scores = fit.step_resident(X.data_ptr(), ev[k])
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    total_ms = ev[0][0].elapsed_time(ev[-1][4])
    phase = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in ev]).mean(0)
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        ph = torch.tensor(phase, device=dev, dtype=torch.float64)
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
        phase = ph.cpu().numpy()
    ms_per_step = total_ms / args.steps
    value = G * S / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI (pinned matrix H2D, index sets H2D, scores D2H)
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
        Xh.copy_(X)
        torch.cuda.synchronize()
        X_host = Xh.numpy()
        e2e_steps = max(1, min(args.steps, 3))
        fit.step_e2e(X_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            scores_e2e = fit.step_e2e(X_host)
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": G * S / dt, "unit": "gene*subsample/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(world * (4 * N * (fit.gl_hi - fit.gl_lo) + idx.nbytes + 4 * N)),
               "d2h_bytes_per_step": int(world * (8 * G + 8 * 4 + 16)),
               "what": "phet_set_classes + phet_set_index_sets(host) + phet_run_pipeline(host pinned matrix: H2D of gene "
                       "block b+1 overlapped with normalise/Step 1/Step 3 of block b) + phet_fisher/o_minmax/bin/"
                       "combine(host scores); wall clock, max over ranks"}
        assert np.allclose(scores_e2e, scores, rtol=1e-9, atol=0)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_step1), live CUDA-event duration on its stream
    if world > 1 and args.shard == "genes":
        # the gene-block pipeline is one C call, so its Step-1 share is re-timed alone on the resident slice
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.step1(fit.s_lo, fit.s_hi, fit.prm)
        e0.record()
        for _ in range(3):
            ctx.step1(fit.s_lo, fit.s_hi, fit.prm)
        e1.record()
        torch.cuda.synchronize()
        phase = np.array([float("nan"), e0.elapsed_time(e1) / 3.0, float("nan"), float("nan")])
    geom = ctx.step1_geometry()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    s_local = fit.s_hi - fit.s_lo
    g_local = (fit.gl_hi - fit.gl_lo) if (world > 1 and args.shard == "genes") else G
    n_pad = (N + 3) // 4 * 4
    hbm_bytes = g_local * n_pad * 4 + 16 * g_local + s_local * 2 * m * 4   # gene rows staged once + accumulators + idx
    smem_bytes = g_local * s_local * 2 * m * 4                             # gathered values
    step1_s = phase[1] * 1e-3
    kname = {"pair-per-thread": "k_step1p<float>", "half-warp": "k_step1h<float,E=%d>" % geom["E"],
             "full-warp": "k_step1<float,E=%d>" % geom["E"]}[geom["path"]]
    roofline = {"bound": "hbm", "kernel": "%s (staged=%s)" % (kname, geom["staged"]),
                "achieved": hbm_bytes / step1_s / 1e9, "peak": peak, "unit": "GB/s",
                "frac": hbm_bytes / step1_s / 1e9 / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full`
                # capture (profiles/r1_ncu_full_v6.txt; config 4 on one GPU only)
                "traffic": NCU_STEP1_DRAM_BYTES if (args.workload == "cfg4" and world == 1 and geom["path"] == "pair-per-thread") else None,
                "peak_source": peak_src,
                "launch_ms": phase[1], "algorithmic_hbm_bytes_per_launch": int(hbm_bytes),
                "smem_gather_gbs": smem_bytes / step1_s / 1e9,
                "smem_gather_peak_gbs": 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9,
                "note": "the Step-1 kernel is not HBM bound: each class vector is staged once per gene (traffic == "
                        "algorithmic bytes) and then serves ~1000 samples from SMEM; ncu: issue slots 61 %, ALU pipe 47 %, "
                        "SMEM wavefronts 44 % (profiles/r1_ncu_full_v6.txt)",
                "grid": geom["grid"], "block": geom["block"], "smem_per_cta": geom["smem"]}
    cpu = None
    if not args.no_cpu:
        g_sub = min(G, 1024)
        cols = X[:, :g_sub].contiguous().cpu().numpy()
        cpu = cpu_reference_sample(wl, cols, ctrl, case, idx)
    line = {"metric": "PHet fit gene*subsample statistics per second", "value": value, "unit": "gene*subsample/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 statistics over f32-staged z-scores",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "phases_ms": {"normalize_transpose": phase[0], "step1": phase[1], "step3": phase[2], "finalize": phase[3]},
            "wall_ms_per_step": t_wall * 1e3 / args.steps, "bin_counts": [int(c) for c in fit.counts],
            "score_sum": float(np.sum(scores))}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_reference_arm(args, wl, config, m):
    """--impl reference: the reference's CPU implementation of the path (the oracle port; the
    Python reference itself cannot travel to the GPU box), single-threaded like the reference,
    each step a bounded sample of the same workload."""
    N, G, S = wl["N"], wl["G"], wl["S"]
    g_sub = min(G, 1024)
    rng = np.random.default_rng(0)
    X = rng.gamma(2.0, 1.0, (N, g_sub)).astype(np.float32)
    k = max(g_sub // 20, 1)
    X[N // 2:, :k] += 1.0
    mu = X[:, k:2 * k].mean(0)
    X[N // 2:, k:2 * k] = (X[N // 2:, k:2 * k] - mu) * 2 + mu
    ctrl, case, idx = draw_index_sets(N, min(S, 64), m)
    res = None
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_reference_sample(wl, X, ctrl, case, idx, g3=8, s_sub=min(S, 24))
        if i >= args.warmup:
            times.append(res["fit_seconds_extrapolated"])
    t_fit = float(np.mean(times))
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
