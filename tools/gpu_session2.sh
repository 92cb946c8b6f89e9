#!/bin/bash
mkdir -p gpurun_out
S=${1:-s2}
if ! timeout 180 python __graft_entry__.py smoke > gpurun_out/${S}_smoke.log 2>&1; then echo SMOKE FAILED; tail -20 gpurun_out/${S}_smoke.log; exit 1; fi
tail -1 gpurun_out/${S}_smoke.log
timeout 200 tools/_bin/lonewarp > gpurun_out/${S}_lonewarp.log 2>&1; cat gpurun_out/${S}_lonewarp.log
: > gpurun_out/${S}_level_probe.jsonl
INF_COOP_MAX=0 timeout 300 python tools/level_probe.py k2 >> gpurun_out/${S}_level_probe.jsonl 2>gpurun_out/${S}_level_probe.err
INF_COOP_MAX=1000000000 timeout 300 python tools/level_probe.py coop >> gpurun_out/${S}_level_probe.jsonl 2>>gpurun_out/${S}_level_probe.err
cat gpurun_out/${S}_level_probe.jsonl
: > gpurun_out/${S}_tree_probe.jsonl
timeout 300 python tools/tree_probe.py default >> gpurun_out/${S}_tree_probe.jsonl 2>gpurun_out/${S}_tree_probe.err
INF_NO_PDL=1 timeout 300 python tools/tree_probe.py nopdl >> gpurun_out/${S}_tree_probe.jsonl 2>>gpurun_out/${S}_tree_probe.err
for c in 32768 65536; do
  INF_COOP_MAX=$c timeout 300 python tools/tree_probe.py coop$c >> gpurun_out/${S}_tree_probe.jsonl 2>>gpurun_out/${S}_tree_probe.err
done
cat gpurun_out/${S}_tree_probe.jsonl
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_multi.py -x -q ) > gpurun_out/${S}_pytest.log 2>&1
tail -5 gpurun_out/${S}_pytest.log
