mkdir -p gpurun_out/s5
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s5/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s5/pytest.log
timeout 600 python bench.py --workload cfg5r8 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/s5/cfg5r8.json 2> gpurun_out/s5/cfg5r8.err; echo "cfg5r8 rc=$?"
( time timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-e2e --no-cpu ) > gpurun_out/s5/cfg5_full.json 2> gpurun_out/s5/cfg5_full.err; echo "cfg5 rc=$?"; tail -4 gpurun_out/s5/cfg5_full.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step1p' -c 1 -o gpurun_out/s5/step1p_big python bench.py --workload cfg5r8 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/s5/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/s5/step1p_big.ncu-rep --page raw --csv > gpurun_out/s5/step1p_big_raw.csv 2>/dev/null
ls -la gpurun_out/s5
python - <<'PY'
import json
for f in ("cfg5r8","cfg5_full"):
    try:
        d=json.load(open("gpurun_out/s5/%s.json"%f)); print(f, d["ms_per_step"], d["phases_ms"], d["roofline"]["block"], d["bin_counts"])
    except Exception as e: print(f, "failed", e)
PY
