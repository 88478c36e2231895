#!/bin/bash
# round 2, first GPU pass: host CPU facts, the new replay tests, then the whole GPU suite
mkdir -p gpurun_out/r2a
lscpu > gpurun_out/r2a/lscpu.txt; nproc >> gpurun_out/r2a/lscpu.txt; free -g >> gpurun_out/r2a/lscpu.txt
nvidia-smi -L >> gpurun_out/r2a/lscpu.txt
python -m pytest tests/test_gpu_replay.py -x -q -m gpu > gpurun_out/r2a/replay_tests.log 2>&1
echo "replay rc=$?" >> gpurun_out/r2a/replay_tests.log
tail -15 gpurun_out/r2a/replay_tests.log
python -m pytest tests -q -m gpu > gpurun_out/r2a/all_tests.log 2>&1
echo "all rc=$?" >> gpurun_out/r2a/all_tests.log
tail -15 gpurun_out/r2a/all_tests.log
python /root/repo/scripts/mt_speed.py > gpurun_out/r2a/mt_speed.txt 2>&1
cat gpurun_out/r2a/mt_speed.txt
