#!/bin/bash
# multi-GPU call: tools/multigpu_check.py (NCCL parity of every partitioned path) + bench.py at N GPUs.  usage: gpu_call_multi.sh <tag> <N> [bench args]
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/multigpu_check.py > gpurun_out/${tag}_mgc.log 2>&1; echo "multigpu_check rc=$?"
grep -E "^\[ok\]|failed|Error|error" gpurun_out/${tag}_mgc.log | head -20
timeout 1200 $TR --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err; echo "bench rc=$?"
tail -3 gpurun_out/${tag}_bench${n}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench${n}.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e'): print(k, d[k])
    print(json.dumps(d.get('spmm_partitioned'))[:3000])
    print(json.dumps(d.get('spmm_c5_full_size'))[:3000])
    for k in ('build_c3_n4','build_c4_n5'): print(k, json.dumps(d.get(k))[:1500])
except Exception as e: print('parse failed', e)
PY
