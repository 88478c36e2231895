#!/bin/bash
# host side of the 8-GPU box: cores, NUMA, and 8 concurrent replays with different thread budgets
mkdir -p gpurun_out/r2p8
{
echo "nproc=$(nproc)"; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|L3|L2" ; cat /sys/fs/cgroup/cpu.max 2>/dev/null; taskset -p $$ | cut -c1-120
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
for cfg in "4 4" "2 2" "2 1" "1 1" "0 1"; do
  set -- $cfg
  for nproc_ in 1 8; do
    echo "== generators=$1 workers=$2 concurrent=$nproc_ (SKIP: count-only + redo without stores)"
    for i in $(seq 1 $nproc_); do SKIP=1 PHET_REPLAY_GENERATORS=$1 /tmp/parbench $2 par | tail -1 & done; wait
  done
done
echo "== generators=4 workers=4 concurrent=8 storing"
for i in $(seq 1 8); do PHET_REPLAY_GENERATORS=4 /tmp/parbench 4 par | tail -1 & done; wait
} 2>&1 | tee gpurun_out/r2p8/probe.txt
