#!/bin/bash
# round 2, session 7: full GPU suite, default bench, ncu launch list + full capture of every kernel shape, sweep
mkdir -p gpurun_out
S=${1:-s7}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${S}_pytest.log 2>&1; tail -3 gpurun_out/${S}_pytest.log
timeout 600 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; echo bench rc $?
timeout 900 python tools/sweep.py --out gpurun_out/r02_sweep_n1.json > gpurun_out/${S}_sweep.log 2>&1; echo sweep rc $?; tail -22 gpurun_out/${S}_sweep.log
timeout 300 python tools/ncu_probe.py > gpurun_out/${S}_ncu_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hash_batch|tree_level|_leaf_kernel' \
    -o gpurun_out/r02_prof python tools/ncu_probe.py > gpurun_out/${S}_ncu_full.log 2>&1; echo ncu full rc $?
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${S}_bench_short.json 2> gpurun_out/${S}_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${S}_ncu_list.log 2>&1; echo ncu list rc $?
ls -la gpurun_out | tail -12
