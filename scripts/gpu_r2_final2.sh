#!/bin/bash
# last pass of round 2 on the final build: full GPU suite, smoke, default bench line, config-5 rank slice + its ncu summary
O=gpurun_out/r2final2; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/all_tests.log 2>&1; tail -2 $O/all_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
timeout 900 python bench.py --workload cfg5r8 --steps 3 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg5r8.json 2> $O/bench_cfg5r8.err
python - <<'PY'
import json
for f in ('bench_n1','bench_cfg5r8'):
    d=json.loads([l for l in open('gpurun_out/r2final2/%s.json'%f) if l.startswith('{')][-1])
    print(f, round(d['ms_per_step'],3), d['phases_ms'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), d['bin_counts'])
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_step1p' -c 1 -o $O/cfg5a python bench.py --workload cfg5r8 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f5a.log 2>&1; echo "ncu rc=$?"
ncu -i $O/cfg5a.ncu-rep --page raw --csv > $O/cfg5a_raw.csv 2>/dev/null; python profiles/summarize_ncu.py $O/cfg5a_raw.csv > $O/ncu_step1p_big.txt; grep -E "kernel|time_duration|issue_active" $O/ncu_step1p_big.txt
rm -f $O/*.ncu-rep
