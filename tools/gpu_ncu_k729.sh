#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spconv_tc_kernel -s 55 -c 6 \
    -f -o gpurun_out/spconv_tc_head python tools/ncu_step.py > gpurun_out/ncu_head.log 2>&1
tail -4 gpurun_out/ncu_head.log
