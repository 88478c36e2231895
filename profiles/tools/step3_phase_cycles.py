"""Per-phase cycle counts of k_step3p as seen by thread 0 of every CTA (env PHET_PROFILE=1 turns the in-kernel
clock64 marks on), for the device bijection and for the reference's permutations read from resident tables.
usage (GPU box): python profiles/tools/step3_phase_cycles.py"""
import os, sys
os.environ["PHET_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from phet_b200 import _capi, hostmath
wl = bench.WORKLOADS["cfg4"]; N, G, S = wl["N"], wl["G"], wl["S"]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
X = bench.synth_matrix_device(N, G, dev)
m = hostmath.subsample_size(N // 2, N - N // 2, "sqrt")
names = ["stage0", "phase1+merge", "stage1", "phase2", "sync", "bounds+flag", "exact", "final sync"]
for mode in ("device", "reference"):
    ctx = _capi.Context(0, stream.cuda_stream)
    ctrl, case, idx, mt = bench.draw_index_sets(N, S, m)
    fit = bench.Fit(ctx, wl, ctrl, case, idx, 0, 1, None, perms=mode)
    fit.setup_resident(X.data_ptr(), mt)
    for _ in range(2): fit.step_resident(X.data_ptr())
    ctx.write(_capi.BUF_PROFILE, np.zeros(16, dtype=np.uint64))
    fit.step_resident(X.data_ptr())
    p = ctx.read(_capi.BUF_PROFILE, np.uint64, (16,)).astype(np.float64)
    tot = p[:8].sum()
    print("== permutations: %s (%s)" % (mode, ctx.step3_path()))
    for n, v in zip(names, p[:8]): print("%-14s %8.0f cycles/gene  %5.1f %%" % (n, v / G, 100 * v / tot))
    print("total per gene %.0f cycles" % (tot / G))
    ctx.close()
