#!/bin/bash
mkdir -p gpurun_out/r2c
python -m pytest tests -x -q -m gpu > gpurun_out/r2c/all_tests.log 2>&1; tail -4 gpurun_out/r2c/all_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c/bench_n1.json 2> gpurun_out/r2c/bench_n1.err; tail -c 3000 gpurun_out/r2c/bench_n1.json; tail -5 gpurun_out/r2c/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c/bench_ref.json 2> gpurun_out/r2c/bench_ref.err; tail -c 1500 gpurun_out/r2c/bench_ref.json; tail -3 gpurun_out/r2c/bench_ref.err
