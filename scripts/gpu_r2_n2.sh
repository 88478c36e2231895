#!/bin/bash
mkdir -p gpurun_out/r2n2
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/r2n2/dist_tests.log 2>&1; tail -3 gpurun_out/r2n2/dist_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n2/bench_n2.json 2> gpurun_out/r2n2/bench_n2.err
tail -5 gpurun_out/r2n2/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n2/bench_n2.json') if l.startswith('{')][-1])
print('value ms', d['ms_per_step'], d['phases_ms'], d['bin_counts'], 'e2e', d['e2e'] and (d['e2e']['ms_per_step'], d['e2e'].get('equals_one_rank_fit_bitwise')), 'dev', d['value_device_perm'] and d['value_device_perm']['ms_per_step'])
PY
