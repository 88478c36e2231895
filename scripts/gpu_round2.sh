mkdir -p gpurun_out/s13
( time timeout 900 python -m pytest tests/test_gpu_multiclass.py -m gpu -q ) > gpurun_out/s13/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s13/pytest_multi.log
timeout 900 python scripts/bench_multiclass.py > gpurun_out/s13/bench_multi.jsonl 2> gpurun_out/s13/bench_multi.err; echo "bench rc=$?"; cat gpurun_out/s13/bench_multi.jsonl
timeout 600 python scripts/multi_fit_once.py > gpurun_out/s13/fit_once.log 2>&1; echo "fit rc=$?"; tail -1 gpurun_out/s13/fit_once.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s13/launches_multi.csv python scripts/multi_fit_once.py > gpurun_out/s13/ncu_l.log 2>&1; echo "ncu launches rc=$?"
python profiles/summarize_launches.py gpurun_out/s13/launches_multi.csv | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_kruskal$' -c 1 -o gpurun_out/s13/kruskal python scripts/multi_fit_once.py 4000 2960 1000 kruskal > gpurun_out/s13/ncu_k.log 2>&1; echo "ncu kruskal rc=$?"
ncu -i gpurun_out/s13/kruskal.ncu-rep --page raw --csv > gpurun_out/s13/kruskal_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/s13/kruskal_raw.csv | head -36
