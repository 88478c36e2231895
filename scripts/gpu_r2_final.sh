#!/bin/bash
# Round-2 evidence on one B200: full GPU suite, default bench line, reference arm, ncu launch list and --set full
# summaries of the hot kernels (config 4 and the config-5 rank slice), phase cycles, multi-class bench.
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/all_tests.log 2>&1; tail -3 $O/all_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -c 600 $O/bench_reference_arm.json
timeout 900 python bench.py --workload cfg5r8 --steps 3 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg5r8.json 2> $O/bench_cfg5r8.err
timeout 900 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg3.json 2> $O/bench_cfg3.err
python - <<'PY'
import json
for f in ('bench_n1','bench_cfg5r8','bench_cfg3'):
    try:
        d=json.loads([l for l in open('gpurun_out/r2final/%s.json'%f) if l.startswith('{')][-1])
        print(f, round(d['ms_per_step'],3), d['phases_ms'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4))
    except Exception as e: print(f, 'failed', e)
PY
# launch list (a number printed under ncu is not a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/ncu_l.log 2>&1; echo "ncu launches rc=$?"
python profiles/summarize_launches.py $O/launches.csv > $O/launches.txt; head -16 $O/launches.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_step1p|k_step3p|k_normalize_transpose|k_colstats_partial' -c 4 -o $O/cfg4 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f.log 2>&1; echo "ncu full cfg4 rc=$?"
ncu -i $O/cfg4.ncu-rep --page raw --csv > $O/cfg4_raw.csv 2>/dev/null
python profiles/summarize_ncu.py $O/cfg4_raw.csv > $O/ncu_full_cfg4.txt; grep -E "kernel|time_duration|dram__bytes_(read|write).sum |issue_active|wavefronts_mem_shared.sum " $O/ncu_full_cfg4.txt | head -30
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_step1p' -c 1 -o $O/cfg5a python bench.py --workload cfg5r8 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f5a.log 2>&1; echo "ncu full cfg5r8 step1p rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_step3s' -c 1 -o $O/cfg5b python bench.py --workload cfg5r8 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f5b.log 2>&1; echo "ncu full cfg5r8 step3s rc=$?"
ncu -i $O/cfg5a.ncu-rep --page raw --csv > $O/cfg5a_raw.csv 2>/dev/null; ncu -i $O/cfg5b.ncu-rep --page raw --csv > $O/cfg5b_raw.csv 2>/dev/null
( python profiles/summarize_ncu.py $O/cfg5a_raw.csv; python profiles/summarize_ncu.py $O/cfg5b_raw.csv ) > $O/ncu_full_cfg5r8.txt; grep -E "kernel|time_duration|dram__bytes_(read|write).sum |issue_active" $O/ncu_full_cfg5r8.txt | head -12
rm -f $O/*.ncu-rep
python profiles/tools/step1_phase_cycles.py > $O/step1_phase_cycles.txt 2>&1; tail -9 $O/step1_phase_cycles.txt
python profiles/tools/step3_phase_cycles.py > $O/step3_phase_cycles.txt 2>&1; tail -3 $O/step3_phase_cycles.txt
timeout 600 python scripts/bench_multiclass.py 10000 30000 1000 > $O/multiclass.jsonl 2> $O/multiclass.err; tail -1 $O/multiclass.jsonl | cut -c1-400
