#!/bin/bash
# host-only timing of the np.random replay on the GPU box's CPU (no GPU work)
mkdir -p gpurun_out/r2cpu
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
for g in 0 2 3 4 6; do for w in 3 4 6; do echo "generators=$g workers=$w"; PHET_REPLAY_GENERATORS=$g /tmp/parbench $w | tail -3 | head -1; done; done > gpurun_out/r2cpu/parbench_ring.txt 2>&1
PHET_REPLAY_GENERATORS=4 /tmp/parbench 4 | tail -2 >> gpurun_out/r2cpu/parbench_ring.txt
cat gpurun_out/r2cpu/parbench_ring.txt
