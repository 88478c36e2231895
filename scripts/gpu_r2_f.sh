#!/bin/bash
mkdir -p gpurun_out/r2f
echo "== new"; timeout 600 python profiles/tools/step1_phase_cycles.py 2>&1 | tail -12 | tee gpurun_out/r2f/phases_new.txt
if [ -f phet_b200/libphet_b200_old.so ]; then
cp phet_b200/libphet_b200.so /tmp/lib_default.so; cp phet_b200/libphet_b200_old.so phet_b200/libphet_b200.so
echo "== old"; timeout 600 python profiles/tools/step1_phase_cycles.py 2>&1 | tail -12 | tee gpurun_out/r2f/phases_old.txt
cp /tmp/lib_default.so phet_b200/libphet_b200.so
fi
