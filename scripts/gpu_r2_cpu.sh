#!/bin/bash
# host-only timing of the np.random replay on the GPU box's CPU (no GPU work)
mkdir -p gpurun_out/r2cpu
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
for pump in 0 3 5 8; do for w in 2 4 6 10; do echo "pump=$pump workers=$w"; PHET_MT_PUMP=$pump /tmp/parbench $w | tail -3; done; done > gpurun_out/r2cpu/parbench.txt 2>&1
cat gpurun_out/r2cpu/parbench.txt
