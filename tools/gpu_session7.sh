#!/bin/bash
mkdir -p gpurun_out
IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_default.log 2>&1; tail -7 gpurun_out/probe_conv_default.log
timeout 200 python tools/probe_attn.py time > gpurun_out/probe_attn_time.log 2>&1; tail -12 gpurun_out/probe_attn_time.log
ONLY=0,1 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_umma -s 6 -c 1 -o gpurun_out/conv_fwd32 -f python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; tail -3 gpurun_out/ncu_conv.log
ONLY=1 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_umma -s 6 -c 1 -o gpurun_out/conv_fwd64 -f python tools/probe_conv.py > gpurun_out/ncu_conv64.log 2>&1; tail -3 gpurun_out/ncu_conv64.log
ls -la gpurun_out/*.ncu-rep
