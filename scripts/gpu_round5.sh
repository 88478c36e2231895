mkdir -p gpurun_out/s17
( time timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q ) > gpurun_out/s17/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s17/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s17/bench_n2.json 2> gpurun_out/s17/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s17/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['bin_counts'], d['config'].get('parallelism'))
PY
tail -3 gpurun_out/s17/bench_n2.err
