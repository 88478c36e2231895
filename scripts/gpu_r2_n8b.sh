#!/bin/bash
mkdir -p gpurun_out/r2n8
echo "nproc=$(nproc)"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2n8/bench_cfg4_n8b.json 2> gpurun_out/r2n8/bench_cfg4_n8b.err
tail -3 gpurun_out/r2n8/bench_cfg4_n8b.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n8/bench_cfg4_n8b.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],3), d['phases_ms'], 'e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],1), d['e2e'].get('equals_one_rank_fit_bitwise'), d['e2e']['timings_rank0']), 'e2e_dev', d['e2e_device_perm'] and round(d['e2e_device_perm']['ms_per_step'],1))
PY
