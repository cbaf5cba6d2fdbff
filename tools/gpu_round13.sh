#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_ref_ops.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/stage_times.py --conv tc 2>&1 | sed -n 3,8p
python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
