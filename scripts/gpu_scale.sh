mkdir -p gpurun_out/s20
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/s20/bench_n$n.json 2> gpurun_out/s20/bench_n$n.err; echo "bench n=$n rc=$?"
done
python - <<'PY'
import json
for n in (8,4,2):
    for l in open('gpurun_out/s20/bench_n%d.json'%n):
        if l.startswith('{'):
            d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['bin_counts'], d['config'].get('parallelism'))
PY
