mkdir -p gpurun_out/s12
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s12/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s12/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/s12/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s12/smoke.log
bash scripts/gpu_multi_bench.sh
python profiles/summarize_ncu.py gpurun_out/s11/anova_raw.csv | head -40
timeout 900 python bench.py > gpurun_out/s12/bench_v10.json 2> gpurun_out/s12/bench_v10.err; echo "bench rc=$?"; head -c 600 gpurun_out/s12/bench_v10.json; echo
timeout 600 python scripts/bench_ingest.py > gpurun_out/s12/bench_ingest.jsonl 2>/dev/null; echo "ingest rc=$?"; cut -c1-400 gpurun_out/s12/bench_ingest.jsonl
