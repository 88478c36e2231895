"""Per-phase cycle counts of k_step1p as seen by thread 0 of every CTA (env PHET_PROFILE=1 turns the
in-kernel clock64 marks on).  usage (GPU box): python profiles/tools/step1_phase_cycles.py"""
import os, sys
os.environ["PHET_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from phet_b200 import _capi, hostmath
wl = bench.WORKLOADS["cfg4"]; N, G, S = wl["N"], wl["G"], wl["S"]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = _capi.Context(0, stream.cuda_stream)
X = bench.synth_matrix_device(N, G, dev)
m = hostmath.subsample_size(N // 2, N - N // 2, "sqrt")
ctrl, case, idx, _ = bench.draw_index_sets(N, S, m)
fit = bench.Fit(ctx, wl, ctrl, case, idx, 0, 1, None, perms="device")
fit.setup_resident(X.data_ptr(), None)
for _ in range(2): fit.step_resident(X.data_ptr())
ctx.write(_capi.BUF_PROFILE, np.zeros(16, dtype=np.uint64))
fit.step_resident(X.data_ptr())
p = ctx.read(_capi.BUF_PROFILE, np.uint64, (16,)).astype(np.float64)
names = ["pass A (gather, moments, histogram)", "locate 4 positions", "bucket-list scan -> records", "records -> element numbers (packs)",
         "candidate gather + sort + pick", "9..16-candidate path, clear histogram", "rescue + ballot", "scratch / combine / p-value"]
passes = p[15]
tot = p[8:15].sum()
for n, v in zip(names, p[8:15]): print("%-40s %9.0f cycles/class-pass  %5.1f %%" % (n, v / passes, 100 * v / tot))
print("total per class pass %.0f cycles (%d passes)" % (tot / passes, passes))
