#!/bin/bash
mkdir -p gpurun_out/r2scan
{
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
for d in 0 1; do for mb in 2 4 8 16; do
echo "== demote=$d ring=${mb}MB"; PHET_RING_DEMOTE=$d PHET_REPLAY_RING_MB=$mb PHET_REPLAY_GENERATORS=3 /tmp/parbench 3 par 2>&1 | tail -2
done; done
} 2>&1 | tee gpurun_out/r2scan/scan4.txt
