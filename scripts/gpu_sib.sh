mkdir -p gpurun_out/s7
( time timeout 900 python -m pytest tests -m gpu -x -q -k "siblings" ) > gpurun_out/s7/pytest_sib.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/s7/pytest_sib.log
timeout 600 python - <<'PY' > gpurun_out/s7/sib_bench.log 2>&1
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import bench
from phet_b200 import _capi, hostmath
N, G = 50000, 30000
X = bench.synth_matrix_device(N, G, torch.device("cuda", 0))
ctx = _capi.Context(0)
ctrl = np.arange(N // 2, dtype=np.int32); case = np.arange(N // 2, N, dtype=np.int32)
ctx.set_classes(ctrl, case)
ctx.load_matrix_device(X.data_ptr(), N, G, _capi.PHET_F32, _capi.NORM_ZSCORE, _capi.PHET_F32)
q3 = []
for n in (N // 2, N - N // 2, N):
    (j0, k0, g0), (j1, k1, g1) = hostmath.quantile_positions(n, (25, 75), dtype=np.float32)
    q3.append(_capi.QuantilePos(j0, k0, j1, k1, g0, g1))
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = ctx.class_stats(q3)
    torch.cuda.synchronize(); print("class_stats cfg4 (50k x 30k f32): %.2f ms" % ((time.perf_counter() - t0) * 1e3))
print(st[:, :3])
PY
cat gpurun_out/s7/sib_bench.log | tail -8
