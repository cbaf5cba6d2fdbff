#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu3.log
tail -30 gpurun_out/pytest_gpu3.log
if grep -q "pytest rc=0" gpurun_out/pytest_gpu3.log; then
timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc3.log 2>&1
head -14 gpurun_out/stage_times_tc3.log; grep -E "K=729|K=125|K=343" gpurun_out/stage_times_tc3.log
CG3D_PAIRS_DEBUG=1 timeout 300 python tools/stage_times.py --conv tc 2>&1 | grep "pairs prof" | tail -3
fi
