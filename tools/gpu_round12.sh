#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -3
for c in 1 0; do echo "== CG3D_COORD_STREAM=$c"; CG3D_COORD_STREAM=$c timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,8p; done
CG3D_COORD_STREAM=1 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
CG3D_COORD_STREAM=0 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
