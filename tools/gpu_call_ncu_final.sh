#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_partitioned_norm_gpu.py -m gpu -q > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_gputest.log
timeout 200 python tools/run_kernels.py count spmm > gpurun_out/${tag}_rk_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"spmm_fan|ngram_count_smem" -s 2 -c 12 -o gpurun_out/${tag}_kernels python tools/run_kernels.py count spmm > gpurun_out/${tag}_rk_ncu.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/${tag}_rk_plain.log | cut -c1-900
