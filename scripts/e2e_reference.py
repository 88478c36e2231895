"""End-to-end fit through the drop-in class with the reference's own random draws (permutation="reference"):
pinned host matrix in, scores out, np.random advanced as the reference would advance it.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS, synth_matrix_device  # noqa: E402
from phet_b200 import PHet  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
N, G, S = wl["N"], wl["G"], wl["S"]
dev = torch.device("cuda", 0)
X = synth_matrix_device(N, G, dev, seed=0)
Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
Xh.copy_(X)
del X
torch.cuda.synchronize()
y = np.zeros(N, dtype=np.int64)
y[N // 2:] = 1
out = {}
for mode in ("reference", "device"):
    times, extra = [], None
    for rep in range(3):
        np.random.seed(12345)
        est = PHet(num_subsamples=S, permutation=mode, distributed=False)
        t0 = time.perf_counter()
        scores = est.fit_predict(Xh.numpy(), y)
        times.append(time.perf_counter() - t0)
        extra = dict(est.timings_)
    out[mode] = {"fit_s": times, "timings": extra, "bin_counts": est.bin_counts_.tolist(), "score_sum": float(scores.sum()),
                 "top10": np.argsort(-scores[:, 0], kind="stable")[:10].tolist()}
print(json.dumps({"workload": wl["name"], "results": out}))
