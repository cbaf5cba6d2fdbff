#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
for d in 0 256; do
echo "== CG3D_TC_DEBUG=$d"
CG3D_TC_DEBUG=$d CG3D_STREAMS=0 timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,14p
done
