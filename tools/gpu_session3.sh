#!/bin/bash
mkdir -p gpurun_out
S=${1:-s3}
if ! timeout 180 python __graft_entry__.py smoke > gpurun_out/${S}_smoke.log 2>&1; then echo SMOKE FAILED; tail -20 gpurun_out/${S}_smoke.log; exit 1; fi
tail -1 gpurun_out/${S}_smoke.log
timeout 200 tools/_bin/lonewarp 2>&1 | grep -i "29\|error" > gpurun_out/${S}_lonewarp.log; cat gpurun_out/${S}_lonewarp.log
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_multi.py -x -q ) > gpurun_out/${S}_pytest.log 2>&1
tail -15 gpurun_out/${S}_pytest.log
