#!/bin/bash
# ncu (+ source-level stall samples) of TS-mode conv launches of one step: SKIP / COUNT select them (27 = the 9^3 class conv)
mkdir -p gpurun_out
for ng in ${NGS:-2}; do
CG3D_TS_NG=$ng timeout 1200 ncu --profile-from-start off --section SpeedOfLight --section SourceCounters --section WarpStateStats --section SchedulerStats \
    --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none --import-source on \
    -k regex:spconv_ts_kernel -s ${SKIP:-27} -c ${COUNT:-1} \
    -f -o gpurun_out/r2_spconv_ts_k729_ng$ng python tools/ncu_step.py > gpurun_out/r2_ncu_ts.log 2>&1
tail -2 gpurun_out/r2_ncu_ts.log
done
ls -la gpurun_out/*.ncu-rep
