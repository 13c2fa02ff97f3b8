#!/bin/bash
# SpMM v3 A/B: register cap / unroll variants built on the box, probe timing for each; then tests + ncu on the default build
tag=${1:-r02c}
mkdir -p gpurun_out
cp protgram-directgcn_b200/libpgb200.so /tmp/default.so
for v in "4 8" "5 8" "6 4" "8 4" "4 4"; do set -- $v
  PGB200_NVCC_FLAGS="-DPG_SPMM_MIN_BLOCKS=$1 -DPG_SPMM_UNROLL1=$2" python protgram-directgcn_b200/build.py --force > /dev/null 2>&1
  python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_mb$1_u$2.log 2>&1; echo "minblocks $1 unroll $2 rc=$?"; tail -1 gpurun_out/${tag}_spmm_mb$1_u$2.log | grep -o "'fanout_fwd': {'ms': [0-9.]*\|'fanin_bwd_operator': {'ms': [0-9.]*\|'fanout_scaled_bwd': {'ms': [0-9.]*"
done
cp /tmp/default.so protgram-directgcn_b200/libpgb200.so
python -m pytest tests/test_gpu_parity.py tests/test_partitioned_norm_gpu.py -m gpu -q -k "spmm or partitioned or halo or model or gradient or tensor_core or layer" > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_gputest.log
python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spmm_fan -s 6 -c 14 -o gpurun_out/${tag}_spmm python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_ncu.log 2>&1
echo "ncu rc=$?"
