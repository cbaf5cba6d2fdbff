#!/bin/bash
mkdir -p gpurun_out
for m in 256 20000 60000; do
echo "== CG3D_MASK_MIN_ROWS=$m"
CG3D_MASK_MIN_ROWS=$m timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,3p
CG3D_MASK_MIN_ROWS=$m CG3D_STREAMS=0 timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,16p
done
