mkdir -p gpurun_out/s11
timeout 900 python scripts/bench_multiclass.py > gpurun_out/s11/bench_multi.jsonl 2> gpurun_out/s11/bench_multi.err; echo "bench rc=$?"; cat gpurun_out/s11/bench_multi.jsonl; tail -5 gpurun_out/s11/bench_multi.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_anova$' -c 1 -o gpurun_out/s11/anova python scripts/bench_multiclass.py 4000 2960 1000 > gpurun_out/s11/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/s11/anova.ncu-rep --page raw --csv > gpurun_out/s11/anova_raw.csv 2>/dev/null
ncu -i gpurun_out/s11/anova.ncu-rep --page source --csv > gpurun_out/s11/anova_source.csv 2>/dev/null
ls -la gpurun_out/s11
