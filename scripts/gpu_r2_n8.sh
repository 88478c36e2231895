#!/bin/bash
# 8 x B200: config 4 at N = 8 and 4 (default bench line), config 5 at N = 8 with the CPU baseline
mkdir -p gpurun_out/r2n8
run() { n=$1; port=$2; out=$3; shift 3
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" ) > gpurun_out/r2n8/$out.json 2> gpurun_out/r2n8/$out.err
  echo "$out rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2n8/$out.err | tail -3 | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2n8/$out.json') if l.startswith('{')][-1])
    print('$out', 'ms', round(d['ms_per_step'],3), d['phases_ms'], d['bin_counts'], 'e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],1), d['e2e'].get('equals_one_rank_fit_bitwise')), 'e2e_dev', d.get('e2e_device_perm') and round(d['e2e_device_perm']['ms_per_step'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))
except Exception as e: print('$out parse failed', e)
PY
}
run 8 29521 bench_cfg4_n8 --steps 20 --warmup 3
run 4 29523 bench_cfg4_n4 --steps 20 --warmup 3
if [ "$1" == "cfg5" ]; then run 8 29522 bench_cfg5_n8 --workload cfg5 --steps 3 --warmup 3 --no-e2e --cpu-at-n; fi
