#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/conv_stress.py > gpurun_out/conv_stress.log 2>&1; rc=$?; echo "conv stress rc=$rc"; tail -16 gpurun_out/conv_stress.log | cut -c1-200
IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_default.log 2>&1; tail -7 gpurun_out/probe_conv_default.log
B2PC_CONV_WS=1 IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_ws.log 2>&1; tail -7 gpurun_out/probe_conv_ws.log
if [ $rc -ne 0 ]; then export B2PC_CONV_V1=1; echo "FALLING BACK TO V1 WGRAD FOR THE REST OF THIS SESSION"; fi
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_default.log 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench_default.log | cut -c1-300; tail -2 gpurun_out/bench.err
B2PC_CONV_WS=1 timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_convws.log 2> gpurun_out/bench2.err; tail -2 gpurun_out/bench2.err
for f in test_gpu_ops test_gpu_scale_parity test_gpu_model; do
  timeout 500 python -m pytest tests/$f.py -q 2>&1 | tail -30 > gpurun_out/pytest_$f.log; echo "$f: $(tail -1 gpurun_out/pytest_$f.log)"
done
B2PC_CONV_WS=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scale_parity.py -q -k "conv or strided" 2>&1 | tail -3
