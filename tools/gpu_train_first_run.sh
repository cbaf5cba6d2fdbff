#!/bin/bash
# First hardware run of the training path (kernels written without a GPU):
#   gpurun --timeout 1500 -- 'bash tools/gpu_train_first_run.sh'
# 1. the provisional parity tests one by one (-rA lists XPASS / XFAIL with reasons; every test runs even if one fails),
# 2. compute-sanitizer memcheck on the kernel-level tests (catches out-of-bounds accesses the comparisons may not show),
# 3. per-kernel device times of a whole training step at the config-4 size (batch 4 x ~50k voxels).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zz_gpu_spconv_backward.py -m gpu -q -rA -p no:cacheprovider \
    > gpurun_out/train_first_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/train_first_pytest.log
grep -E "XPASS|XFAIL|passed|failed|xfailed|xpassed" gpurun_out/train_first_pytest.log | tail -60
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_zz_gpu_spconv_backward.py -m gpu -q -rA \
    -k "wgrad or bn_train or interp or segment_mean or activation or assigner or focal or losses or vote_targets or grouped" \
    > gpurun_out/train_first_memcheck.log 2>&1
grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/train_first_memcheck.log | sort | uniq -c | head -20
timeout 600 python tools/train_times.py --batch 4 --voxels 50000 --steps 3 > gpurun_out/train_first_times.log 2>&1
tail -40 gpurun_out/train_first_times.log
