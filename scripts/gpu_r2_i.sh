#!/bin/bash
mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests/test_gpu_multiclass.py -x -q -m gpu > gpurun_out/r2i/tests.log 2>&1; tail -5 gpurun_out/r2i/tests.log
timeout 600 python scripts/bench_multiclass.py 10000 30000 1000 > gpurun_out/r2i/multiclass.jsonl 2> gpurun_out/r2i/multiclass.err; tail -2 gpurun_out/r2i/multiclass.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2i/multiclass.jsonl') if l.startswith('{')][-1]); print({k:d[k] for k in d if 'ms' in k or 'units' in k})
PY
