#!/bin/bash
# round 2, session 7: default bench, sweep, ncu launch list + full capture of every kernel shape
# (the .ncu-rep is exported to CSV on the box and removed: gpurun_out/ travels back only below 64 MiB)
mkdir -p gpurun_out
S=${1:-s7}
if [ -z "$SKIP_PYTEST" ]; then
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${S}_pytest.log 2>&1; tail -3 gpurun_out/${S}_pytest.log
fi
timeout 600 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; echo bench rc $?
timeout 900 python tools/sweep.py --out gpurun_out/r02_sweep_n1.json > gpurun_out/${S}_sweep.log 2>&1; echo sweep rc $?; tail -4 gpurun_out/${S}_sweep.log
timeout 300 python tools/ncu_probe.py > gpurun_out/${S}_ncu_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hash_batch|tree_level|_leaf_kernel' \
    -o /tmp/r02_prof python tools/ncu_probe.py > gpurun_out/${S}_ncu_full.log 2>&1; echo ncu full rc $?
ncu -i /tmp/r02_prof.ncu-rep --page raw --csv > gpurun_out/r02_prof_raw.csv 2> gpurun_out/${S}_ncu_export.err
ncu -i /tmp/r02_prof.ncu-rep --page details --csv > gpurun_out/r02_prof_details.csv 2>> gpurun_out/${S}_ncu_export.err
ncu -i /tmp/r02_prof.ncu-rep --page source --csv --kernel-name regex:tree_level_coop 2>> gpurun_out/${S}_ncu_export.err | gzip > gpurun_out/r02_prof_source_coop.csv.gz
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${S}_bench_short.json 2> gpurun_out/${S}_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${S}_ncu_list.log 2>&1; echo ncu list rc $?
du -sh gpurun_out; ls -la gpurun_out | tail -14
