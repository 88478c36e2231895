#!/bin/bash
# host-only timing of the np.random replay on the GPU box's CPU (no GPU work)
mkdir -p gpurun_out/r2cpu
cd scripts
g++ -O3 -std=c++17 -o /tmp/calib mt_calib.cpp && /tmp/calib > ../gpurun_out/r2cpu/calib.txt 2>&1
g++ -O3 -std=c++17 -pthread -o /tmp/scanbench mt_scanbench.cpp 2>/dev/null
echo "generic" >> ../gpurun_out/r2cpu/calib.txt; /tmp/scanbench >> ../gpurun_out/r2cpu/calib.txt 2>&1
echo "fast16" >> ../gpurun_out/r2cpu/calib.txt; FAST16=1 /tmp/scanbench | head -3 >> ../gpurun_out/r2cpu/calib.txt 2>&1
g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
for g in 3 4; do echo "generators=$g workers=4"; PHET_REPLAY_DEBUG=1 PHET_REPLAY_GENERATORS=$g /tmp/parbench 4 2>&1 | head -4 | tail -2; done >> gpurun_out/r2cpu/calib.txt
cat gpurun_out/r2cpu/calib.txt
