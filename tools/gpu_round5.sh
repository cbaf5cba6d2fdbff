#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu5.log
tail -30 gpurun_out/pytest_gpu5.log
if grep -q "pytest rc=0" gpurun_out/pytest_gpu5.log; then
timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc5.log 2>&1
head -40 gpurun_out/stage_times_tc5.log; grep -E "K=729|K=125|K=343" gpurun_out/stage_times_tc5.log
CG3D_STREAMS=0 timeout 300 python tools/stage_times.py --conv tc 2>&1 | head -9
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err
fi
