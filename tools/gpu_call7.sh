#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
python tools/run_kernels.py c3step > gpurun_out/${tag}_rk_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_c3.csv python tools/run_kernels.py c3step > gpurun_out/${tag}_rk_ncu.log 2>&1
echo "ncu c3 rc=$?"
python bench.py --steps 4 --warmup 3 --no-large --no-scale --no-cpu-baseline > gpurun_out/${tag}_b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --steps 4 --warmup 3 --no-large --no-scale --no-cpu-baseline > gpurun_out/${tag}_b_ncu.log 2>&1
echo "ncu c2 rc=$?"
