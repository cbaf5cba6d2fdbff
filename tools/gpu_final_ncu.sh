#!/bin/bash
# part B: ncu --set full of the 56 backbone sparse-conv launches of one step; only the summaries travel back (the .ncu-rep
# of 56 launches exceeds gpurun's 64 MiB return limit)
mkdir -p gpurun_out /tmp/ncu
timeout 1500 ncu --profile-from-start off --set full --clock-control none -k regex:spconv_.*_kernel -c 56 \
    -f -o /tmp/ncu/final_spconv_backbone python tools/ncu_step.py > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_ncu_full.log
python tools/ncu_summary.py /tmp/ncu/final_spconv_backbone.ncu-rep gpurun_out/final_ncu_spconv_backbone --traffic-json gpurun_out/final_spconv_traffic.json
cat gpurun_out/final_spconv_traffic.json
ls -la /tmp/ncu gpurun_out
