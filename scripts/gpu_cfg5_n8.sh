mkdir -p gpurun_out/s6
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/s6/cfg5_n8.json 2> gpurun_out/s6/cfg5_n8.err; echo "rc=$?"
tail -5 gpurun_out/s6/cfg5_n8.err; cat gpurun_out/s6/cfg5_n8.json | tail -1 | cut -c1-600
