#!/bin/bash
# ncu --set full (+ source-level stall samples) of the TS-mode conv launches of one step
mkdir -p gpurun_out
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spconv_ts_kernel -c 14 \
    -f -o gpurun_out/r2_spconv_ts python tools/ncu_step.py > gpurun_out/r2_ncu_ts.log 2>&1
tail -3 gpurun_out/r2_ncu_ts.log
ls -la gpurun_out/*.ncu-rep
