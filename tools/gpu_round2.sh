#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu2.log
timeout 900 bash tools/conv_probe.sh > gpurun_out/conv_probe.log 2>&1
timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc2.log 2>&1
tail -3 gpurun_out/pytest_gpu2.log; head -12 gpurun_out/stage_times_tc2.log; cat gpurun_out/conv_probe.log
