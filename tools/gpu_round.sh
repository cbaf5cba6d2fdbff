#!/bin/bash
# one gpurun call: parity tests, stage times, bench, ncu launch list and the --set full capture of the conv kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc.log 2>&1
CG3D_TILE_ORDER_BIG=none timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc_bignone.log 2>&1
CG3D_TC_DEBUG=8 timeout 300 python tools/stage_times.py --conv tc > gpurun_out/stage_times_tc_prof.log 2>&1
timeout 600 python bench.py --layers-json gpurun_out/layers_tc.json > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/ncu_launches_tc.csv python tools/ncu_step.py > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spconv_tc_kernel -c 14 \
    -f -o gpurun_out/spconv_tc python tools/ncu_step.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; head -5 gpurun_out/stage_times_tc.log; cat gpurun_out/bench_tc.json | cut -c1-400
