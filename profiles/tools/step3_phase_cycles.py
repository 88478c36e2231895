import os, sys, json
os.environ["PHET_PROFILE"] = "1"
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from phet_b200 import _capi, hostmath
wl = bench.WORKLOADS["cfg4"]; N, G, S = wl["N"], wl["G"], wl["S"]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = _capi.Context(0, stream.cuda_stream)
X = bench.synth_matrix_device(N, G, dev)
m = hostmath.subsample_size(N // 2, N - N // 2, "sqrt")
ctrl, case, idx = bench.draw_index_sets(N, S, m)
fit = bench.Fit(ctx, wl, ctrl, case, idx, 0, 1, None)
fit.setup_resident(); ctx.set_index_sets(idx)
for _ in range(2): fit.step_resident(X.data_ptr())
z = np.zeros(16, dtype=np.uint64); ctx.write(_capi.BUF_PROFILE, z)
fit.step_resident(X.data_ptr())
p = ctx.read(_capi.BUF_PROFILE, np.uint64, (16,)).astype(np.float64)
names = ["stage0", "phase1+merge", "stage1", "phase2", "sync", "bounds+flag", "exact", "final sync"]
tot = p[:8].sum()
for n, v in zip(names, p[:8]): print("%-14s %8.0f cycles/gene  %5.1f %%" % (n, v / G, 100 * v / tot))
print("total per gene %.0f cycles" % (tot / G))
