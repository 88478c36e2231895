#!/bin/bash
mkdir -p gpurun_out/r2scan
{
cd scripts && g++ -O3 -std=c++17 -pthread -o /tmp/scanbench mt_scanbench.cpp 2>/dev/null && g++ -O3 -std=c++17 -pthread -o /tmp/parbench mt_parbench.cpp 2>/dev/null; cd ..
echo "== count-only scan from a resident ring"; FAST16=1 /tmp/scanbench | head -3
for cfg in "3 3" "3 4" "3 6" "4 6" "2 6"; do set -- $cfg; echo "== generators=$1 workers=$2 storing"; PHET_REPLAY_DEBUG=1 PHET_REPLAY_GENERATORS=$1 /tmp/parbench $2 par 2>&1 | tail -2; done
python -m pytest tests/test_mt_replay_cpu.py -x -q 2>&1 | tail -1
for cfg in "3 3" "3 5" "3 6" "4 6"; do
  set -- $cfg
  PHET_REPLAY_GENERATORS=$1 PHET_REPLAY_WORKERS=$2 python scripts/e2e_once.py 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2scan/scan.txt
