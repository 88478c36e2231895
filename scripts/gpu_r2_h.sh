#!/bin/bash
# full GPU suite + default bench line (N=1) + reference arm
mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h/all_tests.log 2>&1; tail -4 gpurun_out/r2h/all_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h/bench_n1.json 2> gpurun_out/r2h/bench_n1.err; tail -3 gpurun_out/r2h/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2h/bench_n1.json') if l.startswith('{')][-1])
print('value ms', d['ms_per_step'], d['phases_ms'], d['bin_counts'], 'e2e', d['e2e'] and d['e2e']['ms_per_step'], 'dev', d['value_device_perm'] and d['value_device_perm']['ms_per_step'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
PY
