mkdir -p gpurun_out/s10
( time timeout 900 python -m pytest tests/test_gpu_multiclass.py -m gpu -q ) > gpurun_out/s10/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/s10/pytest_multi.log
