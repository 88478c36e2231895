set -x
mkdir -p gpurun_out/s4
( time timeout 900 python -m pytest tests -m gpu -x -q -k "global_gather or big_class" ) > gpurun_out/s4/pytest_big.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/s4/pytest_big.log
for cfgv in "512 0" "512 32" "384 0" "448 0" "320 0"; do
  set -- $cfgv
  PHET_STEP1_BIG_THREADS=$1 PHET_STEP1_NBW=$2 timeout 600 python bench.py --workload cfg5r8 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s4/cfg5r8_t$1_nbw$2.json 2> gpurun_out/s4/cfg5r8_t$1_nbw$2.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/s4/cfg5r8_t$1_nbw$2.json"))
print("threads $1 nbw $2:", d["ms_per_step"], d["phases_ms"], d["roofline"]["kernel"], d["roofline"]["block"], d["roofline"]["smem_per_cta"])
PY
done
