#!/bin/bash
mkdir -p gpurun_out
CG3D_STREAMS=0 timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc6.log 2>&1
head -14 gpurun_out/stage_times_tc6.log; grep -E "K=" gpurun_out/stage_times_tc6.log | awk '{print $2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12}' | sort | uniq -c | sort -rn | head -30
