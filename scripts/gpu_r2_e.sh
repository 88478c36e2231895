#!/bin/bash
# k_step1p with the register-pack post-pass: parity tests, bench; then the serial-RMW variant of pass A
mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_replay.py -x -q -m gpu > gpurun_out/r2e/tests.log 2>&1; tail -5 gpurun_out/r2e/tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2e/bench_n1.json 2> gpurun_out/r2e/bench_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2e/bench_n1.json') if l.startswith('{')][-1]); print('default', d['ms_per_step'], d['phases_ms'], d['bin_counts'])
PY
if [ -f phet_b200/libphet_b200_serial.so ]; then
cp phet_b200/libphet_b200.so /tmp/lib_default.so; cp phet_b200/libphet_b200_serial.so phet_b200/libphet_b200.so
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2e/bench_n1_serial.json 2> gpurun_out/r2e/bench_n1_serial.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2e/bench_n1_serial.json') if l.startswith('{')][-1]); print('serial', d['ms_per_step'], d['phases_ms'], d['bin_counts'])
PY
cp /tmp/lib_default.so phet_b200/libphet_b200.so
fi
