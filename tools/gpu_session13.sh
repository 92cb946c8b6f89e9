#!/bin/bash
# round 2, session 13: functional-basis width-3 kernels, constant first S-box, unconverted round 0 -- full GPU suite, probes, bench
mkdir -p gpurun_out
S=${1:-s13}
t0=$(date +%s)
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc $? $(( $(date +%s) - t0 )) s"; tail -4 gpurun_out/${S}_pytest.log
python tools/variant_probe.py > gpurun_out/${S}_variant.log 2>&1; cat gpurun_out/${S}_variant.log | tail -2
python tools/tree_probe.py default > gpurun_out/${S}_tree_probe.jsonl 2> gpurun_out/${S}_tree_probe.err; cat gpurun_out/${S}_tree_probe.jsonl
timeout 900 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; echo "bench rc $?"; tail -3 gpurun_out/${S}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${S}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac_executed'])
print({k:(v.get('ms'),v.get('hashes_per_s'),v.get('messages_per_s'),v.get('frac_of_pipe_bound')) if isinstance(v,dict) else v for k,v in d['roofline']['configs'].items()})
print(d['roofline']['tree_merge']['ms'], d['roofline']['tree_merge']['state_tree_2^20']['ms'], d['bit_exact_tree'], d['bit_exact_sample'])
PY
