#!/bin/bash
mkdir -p gpurun_out
export IMPLS=2 ONLY=4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_umma -s 6 -c 1 -o gpurun_out/conv_deep256 -f python tools/probe_conv.py > gpurun_out/ncu_deep.log 2>&1; tail -2 gpurun_out/ncu_deep.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_ws -s 3 -c 1 -o gpurun_out/wgrad_deep256 -f python tools/probe_conv.py > gpurun_out/ncu_deep_w.log 2>&1; tail -2 gpurun_out/ncu_deep_w.log
unset ONLY
timeout 200 python -m pytest tests/test_gpu_fused.py -q -k cross_entropy 2>&1 | tail -15
