#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
cp protgram-directgcn_b200/libpgb200.so /tmp/default.so
python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_default.log 2>&1; echo "default (single launch, mb6 u4) rc=$?"; tail -1 gpurun_out/${tag}_spmm_default.log | grep -o "'fanout_fwd': {'ms': [0-9.]*\|'fanin_bwd_operator': {'ms': [0-9.]*\|'fanout_scaled_bwd': {'ms': [0-9.]*"
for v in "-DPG_SPMM_HALF_WARP_ROWS=1" "-DPG_SPMM_HALF_WARP_ROWS=1 -DPG_SPMM_MIN_BLOCKS=5" "-DPG_SPMM_MIN_BLOCKS=7"; do
  PGB200_NVCC_FLAGS="$v" python protgram-directgcn_b200/build.py --force > /dev/null 2>&1
  python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_v.log 2>&1; echo "$v rc=$?"; tail -1 gpurun_out/${tag}_spmm_v.log | grep -o "'fanout_fwd': {'ms': [0-9.]*\|'fanin_bwd_operator': {'ms': [0-9.]*\|'fanout_scaled_bwd': {'ms': [0-9.]*"
done
cp /tmp/default.so protgram-directgcn_b200/libpgb200.so
for c in 256 128; do PGB200_SPMM_CHUNK=$c python tools/run_kernels.py spmm > gpurun_out/${tag}_spmm_chunk$c.log 2>&1; echo "chunk $c rc=$?"; tail -1 gpurun_out/${tag}_spmm_chunk$c.log | grep -o "'fanout_fwd': {'ms': [0-9.]*\|'fanin_bwd_operator': {'ms': [0-9.]*"; done
python -m pytest tests -m gpu -q > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_gputest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "bench ref rc=$?"
