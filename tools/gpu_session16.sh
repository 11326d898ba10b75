#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/conv_ta_check.py > gpurun_out/ta_check.log 2>&1; tail -14 gpurun_out/ta_check.log | cut -c1-160
if ! grep -q "OK" gpurun_out/ta_check.log; then echo "TA path wrong: falling back to B2PC_CONV_TMAA=0 for the rest"; export B2PC_CONV_TMAA=0; fi
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale_parity.py -q -k "conv or strided or spconv" 2>&1 | tail -25 > gpurun_out/pytest_ops.log; tail -4 gpurun_out/pytest_ops.log
timeout 300 python tools/conv_stress.py > gpurun_out/conv_stress.log 2>&1; echo "conv stress rc=$?"; tail -2 gpurun_out/conv_stress.log | cut -c1-200
IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_default.log 2>&1; tail -8 gpurun_out/probe_conv_default.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference --torch-profile gpurun_out/torch_profile_step.txt > gpurun_out/bench_default.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
head -24 gpurun_out/torch_profile_step.txt | cut -c1-150
