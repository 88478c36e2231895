mkdir -p gpurun_out/s4
for cfgv in "352 0" "288 0" "256 0" "224 0" "192 0" "288 32" "256 32"; do
  set -- $cfgv
  PHET_STEP1_BIG_THREADS=$1 PHET_STEP1_NBW=$2 timeout 600 python bench.py --workload cfg5r8 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s4/cfg5r8_t$1_nbw$2.json 2> gpurun_out/s4/cfg5r8_t$1_nbw$2.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/s4/cfg5r8_t$1_nbw$2.json"))
print("threads $1 nbw $2:", d["ms_per_step"], d["phases_ms"]["step1"], d["roofline"]["block"], d["roofline"]["smem_per_cta"])
PY
done
