#!/bin/bash
# round-2 GPU session 2 (every step under its own timeout; a hang costs one step, not the session)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 400 python tools/conv_stress.py > gpurun_out/conv_stress.log 2>&1; rc=$?; echo "conv stress rc=$rc"; tail -16 gpurun_out/conv_stress.log
if [ $rc -ne 0 ]; then export B2PC_CONV_V1=1; echo "FALLING BACK TO V1 CONV KERNELS FOR THIS SESSION"; fi
IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_ws.log 2>&1; tail -7 gpurun_out/probe_conv_ws.log
timeout 200 python tools/probe_attn.py time > gpurun_out/probe_attn_bq32.log 2>&1; tail -3 gpurun_out/probe_attn_bq32.log
B2PC_ATTN_BQ=64 timeout 200 python tools/probe_attn.py time > gpurun_out/probe_attn_bq64.log 2>&1; tail -3 gpurun_out/probe_attn_bq64.log
timeout 300 python -m pytest tests/test_gpu_model.py -q -rP -k "backward_all or autocast" 2>&1 | tail -40 > gpurun_out/pytest_model3.log; tail -25 gpurun_out/pytest_model3.log
timeout 600 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -40 > gpurun_out/pytest_fused.log; tail -12 gpurun_out/pytest_fused.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-supplementary --gpu-reference-steps 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
for f in test_gpu_ops test_gpu_scale_parity test_gpu_model; do
  timeout 500 python -m pytest tests/$f.py -q 2>&1 | tail -30 > gpurun_out/pytest_$f.log; echo "$f: $(tail -1 gpurun_out/pytest_$f.log)"
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
