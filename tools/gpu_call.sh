#!/bin/bash
# One gpurun call of the build loop: GPU tests, SpMM probe + ncu capture, bench line.  usage: tools/gpu_call.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${tag}_gputest.log
python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spmm_fan -s 4 -c 8 -o gpurun_out/${tag}_spmm python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_spmm_plain.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
head -c 1500 gpurun_out/${tag}_bench.json
