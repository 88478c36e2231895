set -x
mkdir -p gpurun_out/s3
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s3/pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/s3/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/s3/bench_v8.json 2> gpurun_out/s3/bench_v8.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3/bench_ref_v8.json 2> gpurun_out/s3/bench_ref_v8.err; echo "ref rc=$?"
( time timeout 600 python bench.py --workload cfg5r8 --steps 2 --warmup 1 --no-e2e --no-cpu ) > gpurun_out/s3/bench_cfg5r8.json 2> gpurun_out/s3/bench_cfg5r8.err; echo "cfg5r8 rc=$?"
tail -3 gpurun_out/s3/pytest.log; cat gpurun_out/s3/smoke.log | tail -2; tail -5 gpurun_out/s3/bench_cfg5r8.err
