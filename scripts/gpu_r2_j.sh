#!/bin/bash
mkdir -p gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_replay.py -x -q -m gpu > gpurun_out/r2j/tests.log 2>&1; tail -3 gpurun_out/r2j/tests.log
for w in cfg5r8 cfg4; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2j/bench_$w.json 2> gpurun_out/r2j/bench_$w.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2j/bench_$w.json') if l.startswith('{')][-1]); print('$w', d['ms_per_step'], d['phases_ms'], d['bin_counts'])
PY
done
