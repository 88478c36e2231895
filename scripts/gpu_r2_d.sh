#!/bin/bash
# scatter Step 3: parity tests, then bench at config 4 (both Step-3 formulations) and the config-5 rank slice
mkdir -p gpurun_out/r2d
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_replay.py tests/test_gpu_multiclass.py -x -q -m gpu > gpurun_out/r2d/tests.log 2>&1; tail -15 gpurun_out/r2d/tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2d/bench_n1.json 2> gpurun_out/r2d/bench_n1.err; tail -c 3000 gpurun_out/r2d/bench_n1.json; tail -5 gpurun_out/r2d/bench_n1.err
PHET_STEP3_IMPL=prune timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2d/bench_n1_prune.json 2> gpurun_out/r2d/bench_n1_prune.err; tail -c 1500 gpurun_out/r2d/bench_n1_prune.json
timeout 600 python bench.py --workload cfg5r8 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2d/bench_cfg5r8.json 2> gpurun_out/r2d/bench_cfg5r8.err; tail -c 2500 gpurun_out/r2d/bench_cfg5r8.json; tail -5 gpurun_out/r2d/bench_cfg5r8.err
