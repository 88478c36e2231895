#!/bin/bash
# ncu --set full of the Step-1 kernel (device-perm bench: fewer kernels before it) + source page
mkdir -p gpurun_out/r2g
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step1p' -c 1 -o gpurun_out/r2g/step1p python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --permutations device > gpurun_out/r2g/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2g/step1p.ncu-rep --page raw --csv > gpurun_out/r2g/step1p_raw.csv 2>/dev/null
ncu -i gpurun_out/r2g/step1p.ncu-rep --page source --csv > gpurun_out/r2g/step1p_source.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r2g/step1p_raw.csv | head -40
rm -f gpurun_out/r2g/step1p.ncu-rep
