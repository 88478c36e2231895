#!/bin/bash
mkdir -p gpurun_out/r2b
python -m pytest tests/test_gpu_replay.py -x -q -m gpu > gpurun_out/r2b/replay_tests.log 2>&1; tail -3 gpurun_out/r2b/replay_tests.log
python scripts/e2e_reference.py cfg4 > gpurun_out/r2b/e2e_reference_cfg4.json 2> gpurun_out/r2b/e2e_reference_cfg4.err; cat gpurun_out/r2b/e2e_reference_cfg4.json; tail -5 gpurun_out/r2b/e2e_reference_cfg4.err
