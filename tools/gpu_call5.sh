#!/bin/bash
tag=${1:-r02f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${tag}_gputest.log
python tools/run_kernels.py spmm c3step > gpurun_out/${tag}_rk_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python tools/run_kernels.py c3step > gpurun_out/${tag}_rk_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${tag}_rk_plain.log | cut -c1-1500
