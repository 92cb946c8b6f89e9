#!/bin/bash
# round 2, session 5: rotated role layout in the cooperative kernel, 64-thread blocks for mid-size levels,
# small batches through the cooperative kernel, one-launch zero chains, ramped pageable pipeline
mkdir -p gpurun_out
S=${1:-s5}
if ! timeout 180 python __graft_entry__.py smoke > gpurun_out/${S}_smoke.log 2>&1; then echo SMOKE FAILED; tail -20 gpurun_out/${S}_smoke.log; exit 1; fi
tail -1 gpurun_out/${S}_smoke.log
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "small_and_large or zero_tables or pageable or all_widths or dense_kernel or tree_merge_equals or domain_tag or 2_16" ) > gpurun_out/${S}_pytest_quick.log 2>&1
tail -4 gpurun_out/${S}_pytest_quick.log
{
python tools/level_probe.py default
INF_NO_SMALL_BLOCKS=1 INF_COOP_MAX=0 python tools/level_probe.py k2_128
INF_COOP_MAX=0 python tools/level_probe.py k2_smallblocks
INF_COOP_MAX=1000000000 python tools/level_probe.py coop_rotated
} > gpurun_out/${S}_level_probe.jsonl 2> gpurun_out/${S}_level_probe.err
cat gpurun_out/${S}_level_probe.jsonl
{
python tools/tree_probe.py default
INF_NO_SMALL_BLOCKS=1 python tools/tree_probe.py no_small_blocks
INF_COOP_MAX=32768 python tools/tree_probe.py coop32768
} > gpurun_out/${S}_tree_probe.jsonl 2> gpurun_out/${S}_tree_probe.err
cat gpurun_out/${S}_tree_probe.jsonl
