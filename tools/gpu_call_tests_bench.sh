#!/bin/bash
tag=${1:-r02h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${tag}_gputest.log | cut -c1-400
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'])
print('e2e',json.dumps(d['e2e'])[:900])
print('c3',json.dumps(d.get('directgcn_c3'))[:500])
print('api',json.dumps(d.get('e2e_api'))[:600])
print(d['per_step_ms'])
PY
