mkdir -p gpurun_out/s19
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/s19/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s19/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/s19/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s19/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s19/bench_ref.json 2> gpurun_out/s19/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/s19/bench_ref.json
timeout 900 python bench.py > gpurun_out/s19/bench_v12.json 2> gpurun_out/s19/bench_v12.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s19/bench_v12.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['phases_ms'], d['bin_counts'], d['score_sum'], d['cpu_baseline']['value'], d['clocks'])
PY
