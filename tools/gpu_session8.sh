#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_data.py tests/test_gpu_eval_ops.py -q 2>&1 | tail -40 > gpurun_out/pytest_data.log; tail -25 gpurun_out/pytest_data.log
for mb in 1 0; do
  export B2PC_CONV_MB=$mb
  echo "== MB=$mb"
  timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale_parity.py -q -k "conv or strided or spconv" 2>&1 | tail -8 > gpurun_out/pytest_conv_mb$mb.log; tail -3 gpurun_out/pytest_conv_mb$mb.log
  IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_mb$mb.log 2>&1; tail -7 gpurun_out/probe_conv_mb$mb.log
done
unset B2PC_CONV_MB
timeout 300 python tools/conv_stress.py > gpurun_out/conv_stress.log 2>&1; echo "conv stress rc=$?"; tail -4 gpurun_out/conv_stress.log | cut -c1-200
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_default.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
