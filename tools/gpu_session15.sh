#!/bin/bash
# round 2, session 15: history-recurrence partial rounds (widths 2..6) -- A/B against the build without them
# and against two rounds per loop body, full GPU suite, bench, ncu (headline, launch list, every kernel shape)
mkdir -p gpurun_out
S=${1:-s15}
{ python tools/variant_probe.py; INFIMUM_B200_LIB=$PWD/tools/_bin/libinfimum_b200_nohr.so python tools/variant_probe.py;
  INFIMUM_B200_LIB=$PWD/tools/_bin/libinfimum_b200_unr2.so python tools/variant_probe.py; } > gpurun_out/${S}_variant.log 2>&1; grep hash2 gpurun_out/${S}_variant.log
t0=$(date +%s)
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc $? $(( $(date +%s) - t0 )) s"; tail -4 gpurun_out/${S}_pytest.log
python tools/tree_probe.py default > gpurun_out/${S}_tree_probe.jsonl 2> gpurun_out/${S}_tree_probe.err; cat gpurun_out/${S}_tree_probe.jsonl
timeout 600 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; echo "bench rc $?"; tail -3 gpurun_out/${S}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${S}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac_executed'])
print({k:(v.get('ms'),v.get('hashes_per_s'),v.get('messages_per_s'),v.get('frac_of_pipe_bound')) if isinstance(v,dict) else v for k,v in d['roofline']['configs'].items()})
print(d['roofline']['tree_merge']['ms'], d['roofline']['tree_merge']['state_tree_2^20']['ms'], d['bit_exact_tree'], d['bit_exact_sample'])
PY
timeout 200 python tools/ncu_probe.py headline > gpurun_out/${S}_ncu_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'hash_batch_kernel' -o /tmp/r02c_headline \
    python tools/ncu_probe.py headline > gpurun_out/${S}_ncu_headline.log 2>&1; echo ncu headline rc $?
ncu -i /tmp/r02c_headline.ncu-rep --page raw --csv > gpurun_out/r02c_hash2_headline_raw.csv 2> gpurun_out/${S}_ncu_export.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${S}_ncu_list.log 2>&1; echo ncu list rc $?
timeout 200 python tools/ncu_probe.py >> gpurun_out/${S}_ncu_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hash_batch|tree_level|_leaf_kernel' \
    -o /tmp/r02c_prof python tools/ncu_probe.py > gpurun_out/${S}_ncu_full.log 2>&1; echo ncu full rc $?
ncu -i /tmp/r02c_prof.ncu-rep --page raw --csv > gpurun_out/r02c_prof_raw.csv 2>> gpurun_out/${S}_ncu_export.err
ncu -i /tmp/r02c_prof.ncu-rep --page details --csv > gpurun_out/r02c_prof_details.csv 2>> gpurun_out/${S}_ncu_export.err
