#!/bin/bash
# Round-2 evidence run (one B200): parity tests, smoke, the four bench workloads + reference arm, stage times, training kernel
# times, ncu launch list, ncu --set full of the backbone conv launches (+ traffic json) and of the head / training kernels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_final_r2.sh'
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > $O/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/final_pytest_gpu.log
tail -3 $O/final_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/final_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/final_clocks.csv &
SMI=$!
timeout -k 5 600 python bench.py --layers-json $O/final_layers.json > $O/final_bench.json 2> $O/final_bench.err
kill $SMI
cut -c1-400 $O/final_bench.json
timeout -k 5 300 python bench.py --workload sunrgbd --steps 5 --warmup 3 > $O/final_bench_sunrgbd.json 2>> $O/final_bench.err
timeout -k 5 400 python bench.py --workload sweep --steps 5 --warmup 3 > $O/final_bench_sweep.json 2>> $O/final_bench.err
timeout -k 5 300 python bench.py --workload train --steps 5 --warmup 3 > $O/final_bench_train.json 2>> $O/final_bench.err
timeout -k 5 300 python bench.py --workload train_sunrgbd --steps 5 --warmup 3 --no-cpu-baseline > $O/final_bench_train_sunrgbd.json 2>> $O/final_bench.err
timeout -k 5 120 python -m pytest tests/test_gpu_ops.py -m gpu -q -s -k "nms_long_segments" 2>&1 | grep "^NMS\|passed\|failed" > $O/final_nms_paths.log
timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference.json 2>> $O/final_bench.err
for f in sunrgbd sweep train train_sunrgbd reference; do cut -c1-200 $O/final_bench_$f.json; done
timeout -k 5 300 python tools/stage_times.py --conv tc > $O/final_stage_times.log 2>&1
timeout -k 5 300 python tools/train_times.py --batch 4 --voxels 50000 --steps 3 > $O/final_train_times.log 2>&1
timeout -k 5 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/final_ncu_launches.csv python tools/ncu_step.py > $O/final_ncu_launches.log 2>&1
# ncu --set full: the 56 backbone conv launches (summary + DRAM traffic per launch for bench.py's roofline.traffic)
timeout -k 5 1200 ncu --profile-from-start off --set full --clock-control none -k regex:spconv_.*_kernel -c 56 \
    -f -o /tmp/ncu/final_spconv_backbone python tools/ncu_step.py > $O/final_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/final_spconv_backbone.ncu-rep $O/final_ncu_spconv_backbone --traffic-json $O/final_spconv_traffic.json
# the head / RoI conv launches (the 9^3 class conv among them): launches 56.. of the same kernels
timeout -k 5 900 ncu --profile-from-start off --set full --clock-control none -k regex:spconv_.*_kernel -s 56 -c 16 \
    -f -o /tmp/ncu/final_spconv_head python tools/ncu_step.py > $O/final_ncu_head.log 2>&1
python tools/ncu_summary.py /tmp/ncu/final_spconv_head.ncu-rep $O/final_ncu_spconv_head
ls -la /tmp/ncu $O | tail -40
# the selection primitives rewritten at the end of round 2 (one-sweep sort, chained scan, blocked NMS)
timeout -k 5 300 ncu --profile-from-start off --set full --clock-control none -k regex:"rs_onesweep_kernel|rs_hist_kernel|scan_chained_kernel|blocked_nms_kernel" -c 14 \
    -f -o /tmp/ncu/final_select python tools/ncu_step.py > $O/final_ncu_select.log 2>&1
python tools/ncu_summary.py /tmp/ncu/final_select.ncu-rep $O/final_ncu_select
timeout -k 5 120 python tools/timeline.py > $O/final_timeline.log 2>&1
ls -la /tmp/ncu $O | tail -40
