#!/bin/bash
# TS-mode (A operand in TMEM) conv kernel: parity tests, then per-layer timing: default shape choice | NG=2 | NG=4 | off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -15
CG3D_TS_NG=4 timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -3
PROBE_DEBUGS=0 timeout 300 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_on.log 2>&1
CG3D_TS_NG=2 PROBE_DEBUGS=0 timeout 300 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_ng2.log 2>&1
CG3D_TS_NG=4 PROBE_DEBUGS=0 timeout 300 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_ng4.log 2>&1
CG3D_TC_TS=0 PROBE_DEBUGS=0 timeout 300 python tools/conv_probe2.py > gpurun_out/r2_probe_ts_off.log 2>&1
paste gpurun_out/r2_probe_ts_on.log gpurun_out/r2_probe_ts_ng2.log gpurun_out/r2_probe_ts_ng4.log gpurun_out/r2_probe_ts_off.log | awk '{print $2,$3,$4,$6,"| ng2",$13,"| ng4",$20,"| off", $27}' | grep -E "Cin=64|K=125|K=343"
PROBE_MIN_K=729 PROBE_DEBUGS=8 python tools/conv_probe2.py 2>&1 | grep "ts prof\|spconv_tc K" | head -2
