#!/bin/bash
mkdir -p gpurun_out/r2n2
python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/r2n2/dist_tests.log 2>&1; tail -3 gpurun_out/r2n2/dist_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n2/bench_n2.json 2> gpurun_out/r2n2/bench_n2.err
tail -c 2500 gpurun_out/r2n2/bench_n2.json; tail -5 gpurun_out/r2n2/bench_n2.err
