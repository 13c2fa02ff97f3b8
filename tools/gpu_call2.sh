#!/bin/bash
# SpMM v2 check: parity tests of the SpMM / partitioned paths, timing at two long-row chunk sizes, ncu capture
tag=${1:-r02b}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_partitioned_norm_gpu.py -m gpu -q -k "spmm or partitioned or halo or model or gradient or tensor_core or layer" > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${tag}_gputest.log
for c in 512 256 128; do PGB200_SPMM_CHUNK=$c python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_chunk$c.log 2>&1; echo "chunk $c rc=$?"; tail -1 gpurun_out/${tag}_spmm_chunk$c.log; done
python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spmm_fan -s 4 -c 8 -o gpurun_out/${tag}_spmm python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_ncu.log 2>&1
echo "ncu rc=$?"
