#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu4.log
tail -30 gpurun_out/pytest_gpu4.log
if grep -q "pytest rc=0" gpurun_out/pytest_gpu4.log; then
timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc4.log 2>&1
head -16 gpurun_out/stage_times_tc4.log; grep -E "K=729|K=125|K=343| 64->  64" gpurun_out/stage_times_tc4.log
CG3D_TC_STACKED=0 timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc4_nostack.log 2>&1
head -3 gpurun_out/stage_times_tc4_nostack.log; grep -E "K=729|K=125|K=343| 64->  64" gpurun_out/stage_times_tc4_nostack.log
fi
