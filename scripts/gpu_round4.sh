mkdir -p gpurun_out/s16
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s16/launches_v11.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s16/ncu_l.log 2>&1; echo "ncu launches rc=$?"
python profiles/summarize_launches.py gpurun_out/s16/launches_v11.csv | head -24
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step1p|k_step3p|k_normalize_transpose|k_colstats_partial' -c 4 -o gpurun_out/s16/v11 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/s16/ncu_f.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/s16/v11.ncu-rep --page raw --csv > gpurun_out/s16/v11_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/s16/v11_raw.csv | grep -E "kernel|time_duration|dram__bytes|issue_active|warps_active" 
