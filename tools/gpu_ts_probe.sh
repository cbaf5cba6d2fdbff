#!/bin/bash
# TS-mode (A operand in TMEM) conv kernel: parity tests, then per-layer timing: 64- and 128-column tiles | 64 only | off
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -4
PROBE_DEBUGS=0 timeout -k 5 200 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_on.log 2>&1
CG3D_TC_TS=64 PROBE_DEBUGS=0 timeout -k 5 200 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_64.log 2>&1
CG3D_TC_TS=0 PROBE_DEBUGS=0 timeout -k 5 200 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_off.log 2>&1
paste gpurun_out/r2_probe_ts_on.log gpurun_out/r2_probe_ts_64.log gpurun_out/r2_probe_ts_off.log | awk '{print $2,$3,$4,$6,"| ts64",$13,"| off", $20}'
