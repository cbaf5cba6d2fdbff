#!/bin/bash
# TS-mode (A operand in TMEM) conv kernel: parity tests, then per-layer timing (all layers incl. K = 1): default | column tile capped at 128 everywhere | TS off
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -4
export PROBE_MIN_K=1
PROBE_DEBUGS=0 timeout -k 5 200 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_on.log 2>&1
CG3D_TC_NTMAX=128 PROBE_DEBUGS=0 timeout -k 5 200 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_nt128.log 2>&1
CG3D_TC_TS=0 PROBE_DEBUGS=0 timeout -k 5 200 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_off.log 2>&1
paste gpurun_out/r2_probe_ts_on.log gpurun_out/r2_probe_ts_nt128.log gpurun_out/r2_probe_ts_off.log | awk '{print $2,$3,$4,$6,"| NT<=128",$13,"| off", $20}'
