#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu9.log
tail -12 gpurun_out/pytest_gpu9.log
for gmode in tma cpasync; do
echo "== CG3D_TC_GATHER=$gmode"
CG3D_TC_GATHER=$gmode CG3D_STREAMS=0 timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_$gmode.log 2>&1
sed -n 3,14p gpurun_out/stage_times_$gmode.log; grep -E "K=729|K=125|K=343" gpurun_out/stage_times_$gmode.log
done
CG3D_TC_GATHER=tma timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,3p
