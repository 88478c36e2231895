"""One multi-class fit (K = 3) for the ncu launch list: python scripts/multi_fit_once.py [N G S group_test]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from phet_b200 import PHet  # noqa: E402

N, G, S = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (10000, 30000, 1000)
gt = sys.argv[4] if len(sys.argv) > 4 else "anova"
X = bench.synth_matrix_device(N, G, torch.device("cuda", 0)).cpu().numpy()
y = (np.arange(N) * 3 // N).astype(np.int64)
np.random.seed(12345)
est = PHet(num_subsamples=S, device=0, permutation="device", group_test=gt)
scores = est.fit_predict(X, y)
print("fit ok", scores.sum(), est.launches_, est.timings_)
