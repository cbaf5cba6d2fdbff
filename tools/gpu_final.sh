#!/bin/bash
# the round's evidence run, part A: parity tests, bench (with the CPU baseline), reference arm, stage times, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/final_pytest_gpu.log
tail -3 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/final_clocks.csv &
SMI=$!
timeout 900 python bench.py --layers-json gpurun_out/final_layers.json > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
kill $SMI
cut -c1-600 gpurun_out/final_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
cut -c1-300 gpurun_out/final_bench_reference.json
timeout 300 python tools/stage_times.py --conv tc > gpurun_out/final_stage_times.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/final_ncu_launches.csv python tools/ncu_step.py > gpurun_out/final_ncu_launches.log 2>&1
ls -la gpurun_out
