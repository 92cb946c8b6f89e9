#!/bin/bash
# One GPU-box session: tests, microbenchmarks, probes, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
S=${1:-s1}
if ! timeout 180 python __graft_entry__.py smoke > gpurun_out/${S}_smoke.log 2>&1; then echo SMOKE FAILED; tail -20 gpurun_out/${S}_smoke.log; exit 1; fi
tail -2 gpurun_out/${S}_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${S}_pytest.log 2>&1
tail -5 gpurun_out/${S}_pytest.log
timeout 120 tools/_bin/lonewarp > gpurun_out/${S}_lonewarp.log 2>&1
: > gpurun_out/${S}_tree_probe.jsonl
timeout 300 python tools/tree_probe.py default >> gpurun_out/${S}_tree_probe.jsonl 2>gpurun_out/${S}_tree_probe.err
INF_NO_PDL=1 timeout 300 python tools/tree_probe.py nopdl >> gpurun_out/${S}_tree_probe.jsonl 2>>gpurun_out/${S}_tree_probe.err
for c in 0 2048 4096 8192 32768 65536; do
  INF_COOP_MAX=$c timeout 300 python tools/tree_probe.py coop$c >> gpurun_out/${S}_tree_probe.jsonl 2>>gpurun_out/${S}_tree_probe.err
done
cat gpurun_out/${S}_tree_probe.jsonl
cat gpurun_out/${S}_lonewarp.log
timeout 900 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err
tail -c 3000 gpurun_out/${S}_bench.json
