mkdir -p gpurun_out/s15
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/s15/pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^E   *\(+\|array\|\[\)" gpurun_out/s15/pytest.log | tail -15
timeout 300 python __graft_entry__.py smoke > gpurun_out/s15/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s15/smoke.log
timeout 900 python bench.py > gpurun_out/s15/bench_v11.json 2> gpurun_out/s15/bench_v11.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s15/bench_v11.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['phases_ms'], d['bin_counts'], d['score_sum'])
PY
