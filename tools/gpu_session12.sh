#!/bin/bash
# round 2, session 12: validation of HEAD after the container was re-created (full GPU suite, smoke, both bench arms)
mkdir -p gpurun_out
S=${1:-s12}
t0=$(date +%s)
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc $? $(( $(date +%s) - t0 )) s"; tail -14 gpurun_out/${S}_pytest.log
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${S}_smoke.log 2>&1; echo "smoke rc $? $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/${S}_smoke.log
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; echo "bench rc $? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/${S}_bench.err
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${S}_bench_ref.json 2> gpurun_out/${S}_bench_ref.err; echo "ref rc $? $(( $(date +%s) - t0 )) s"
python - <<PY
import json
d=json.loads(open('gpurun_out/${S}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac_executed'])
print({k:(v.get('ms'),v.get('hashes_per_s'),v.get('messages_per_s')) if isinstance(v,dict) else v for k,v in d['roofline']['configs'].items()})
print({k:v for k,v in d['e2e'].items() if isinstance(v,dict)})
print(d['roofline']['tree_merge']['ms'], d['roofline']['tree_merge']['state_tree_2^20']['ms'], d['bit_exact_tree'], d['bit_exact_sample'])
r=json.loads(open('gpurun_out/${S}_bench_ref.json').read().strip().splitlines()[-1]); print(r['value'], r['cpu_baseline']['cores'])
PY
