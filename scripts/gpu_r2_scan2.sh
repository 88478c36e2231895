#!/bin/bash
mkdir -p gpurun_out/r2scan
{
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/scanbench mt_scanbench.cpp 2>/dev/null && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
for pf in 0 512 1024 2048 4096; do
echo "== prefetch $pf: scan from a resident ring / parbench 3+3"; PHET_SCAN_PREFETCH=$pf FAST16=1 /tmp/scanbench | head -3 | tail -2
PHET_SCAN_PREFETCH=$pf PHET_REPLAY_GENERATORS=3 /tmp/parbench 3 par 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2scan/scan2.txt
