"""One process, four end-to-end fits of config 4 through the drop-in class (reference draws); prints the wall times."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from bench import WORKLOADS, synth_matrix_device
from phet_b200 import PHet
wl = WORKLOADS["cfg4"]; N, G, S = wl["N"], wl["G"], wl["S"]
dev = torch.device("cuda", 0)
X = synth_matrix_device(N, G, dev, seed=0)
Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=True); Xh.copy_(X); del X
torch.cuda.synchronize()
y = np.zeros(N, dtype=np.int64); y[N // 2:] = 1
ts = []
for rep in range(4):
    np.random.seed(12345)
    est = PHet(num_subsamples=S, distributed=False)
    t0 = time.perf_counter(); est.fit_predict(Xh.numpy(), y); ts.append(time.perf_counter() - t0)
print(os.environ.get("PHET_REPLAY_GENERATORS"), os.environ.get("PHET_REPLAY_WORKERS"), ["%.3f" % t for t in ts], est.timings_.get("draw_thread_s"))
