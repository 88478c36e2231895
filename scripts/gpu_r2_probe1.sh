#!/bin/bash
mkdir -p gpurun_out/r2p1
{
echo "nproc=$(nproc)"; lscpu | grep -E "Model name|Socket|Core|Thread|L3|L2"
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
for cfg in "4 4" "3 3" "2 2" "2 1" "1 1" "1 2" "3 1" "6 2"; do
  set -- $cfg
  echo "== generators=$1 workers=$2 storing"
  PHET_REPLAY_GENERATORS=$1 /tmp/parbench $2 par | tail -1
done
} 2>&1 | tee gpurun_out/r2p1/probe.txt
