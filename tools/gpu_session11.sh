#!/bin/bash
mkdir -p gpurun_out
export IMPLS=2 ONLY=3,4,6
: > gpurun_out/conv_deep_sweep.log
for nt in 0 128 64; do for fill in 1 2 4; do for ks in 9 27; do
  echo "== NT=$nt FILL=$fill KSPLIT=$ks" >> gpurun_out/conv_deep_sweep.log
  B2PC_CONV_NT=$nt B2PC_CONV_FILL=$fill B2PC_CONV_KSPLIT=$ks timeout 120 python tools/probe_conv.py 2>&1 | tail -3 | cut -c1-140 >> gpurun_out/conv_deep_sweep.log
done; done; done
cat gpurun_out/conv_deep_sweep.log
unset ONLY
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale_parity.py -q -k "conv or strided or spconv" 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_fused.py -q 2>&1 | tail -3
