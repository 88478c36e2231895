mkdir -p gpurun_out/s8
cat > /tmp/sibrun.py <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import bench
from phet_b200 import _capi, hostmath
N, G = 50000, 2960
X = bench.synth_matrix_device(N, G, torch.device("cuda", 0))
ctx = _capi.Context(0)
ctrl = np.arange(N // 2, dtype=np.int32); case = np.arange(N // 2, N, dtype=np.int32)
ctx.set_classes(ctrl, case)
ctx.load_matrix_device(X.data_ptr(), N, G, _capi.PHET_F32, _capi.NORM_ZSCORE, _capi.PHET_F32)
q3 = []
for n in (N // 2, N - N // 2, N):
    (j0, k0, g0), (j1, k1, g1) = hostmath.quantile_positions(n, (25, 75), dtype=np.float32)
    q3.append(_capi.QuantilePos(j0, k0, j1, k1, g0, g1))
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = ctx.class_stats(q3)
    torch.cuda.synchronize(); print("class_stats (50k x 2960 f32): %.2f ms" % ((time.perf_counter() - t0) * 1e3))
PY
timeout 600 python /tmp/sibrun.py > gpurun_out/s8/plain.log 2>&1; cat gpurun_out/s8/plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_class_stats' -c 1 -o gpurun_out/s8/cs python /tmp/sibrun.py > gpurun_out/s8/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/s8/cs.ncu-rep --page raw --csv > gpurun_out/s8/cs_raw.csv 2>/dev/null
ncu -i gpurun_out/s8/cs.ncu-rep --page source --csv > gpurun_out/s8/cs_source.csv 2>/dev/null
ls -la gpurun_out/s8
