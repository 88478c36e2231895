mkdir -p gpurun_out/s9
( time timeout 900 python -m pytest tests/test_gpu_ingest.py -m gpu -x -q ) > gpurun_out/s9/pytest_ingest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/s9/pytest_ingest.log
timeout 600 python scripts/bench_ingest.py > gpurun_out/s9/bench_ingest.jsonl 2> gpurun_out/s9/bench_ingest.err; echo "bench rc=$?"; cat gpurun_out/s9/bench_ingest.jsonl; tail -5 gpurun_out/s9/bench_ingest.err
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli or srbct or errors" ) > gpurun_out/s9/pytest_cli.log 2>&1; echo "pytest cli rc=$?"; tail -5 gpurun_out/s9/pytest_cli.log
