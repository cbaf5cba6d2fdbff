#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_tc.py -m gpu -q 2>&1 | tail -3
for d in 0 512; do echo "== debug=$d"; CG3D_TC_DEBUG=$d CG3D_STREAMS=0 timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,12p; done
timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,3p
