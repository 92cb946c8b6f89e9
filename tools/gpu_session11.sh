#!/bin/bash
# round 2, session 11: wide-block geometry (one 384/512-thread block per SM) chosen per launch by wave fill
mkdir -p gpurun_out
S=${1:-s11}
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${S}_pytest.log 2>&1; tail -3 gpurun_out/${S}_pytest.log
{ python tools/geom_probe.py default; INF_WIDE_BLOCK=128 INF_LEAF_BLOCK=128 python tools/geom_probe.py all128; } > gpurun_out/${S}_geom.jsonl 2> gpurun_out/${S}_geom.err; cat gpurun_out/${S}_geom.jsonl
python tools/tree_probe.py default > gpurun_out/${S}_tree_probe.jsonl 2> gpurun_out/${S}_tree_probe.err; cat gpurun_out/${S}_tree_probe.jsonl
python tools/replay_probe.py default > gpurun_out/${S}_replay_probe.jsonl 2>&1; cat gpurun_out/${S}_replay_probe.jsonl
timeout 600 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; echo bench rc $?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s11_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac_executed'])
print({k:(v.get('ms'),v.get('hashes_per_s'),v.get('messages_per_s')) for k,v in d['roofline']['configs'].items()})
print({k:v for k,v in d['e2e'].items() if isinstance(v,dict)})
print(d['roofline']['tree_merge']['ms'], d['roofline']['tree_merge']['state_tree_2^20']['ms'])
PY
