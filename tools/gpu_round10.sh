#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
CG3D_STREAMS=0 timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc10.log 2>&1
sed -n 3,14p gpurun_out/stage_times_tc10.log; grep -E "K=729|K=125|K=343" gpurun_out/stage_times_tc10.log
timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,3p
