#!/bin/bash
mkdir -p gpurun_out/r2scan
{
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/scanbench mt_scanbench.cpp 2>/dev/null && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
echo "== scan from a resident ring / parbench 3+3"; FAST16=1 /tmp/scanbench | head -3 | tail -2
PHET_REPLAY_GENERATORS=3 /tmp/parbench 3 par 2>&1 | tail -1
python -m pytest tests/test_mt_replay_cpu.py tests/test_gpu_replay.py -x -q -m "gpu or not gpu" 2>&1 | tail -1
python scripts/e2e_once.py 2>&1 | tail -1
PHET_REPLAY_WORKERS=5 python scripts/e2e_once.py 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2scan/scan3.txt
