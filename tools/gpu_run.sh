#!/bin/bash
# usage (under gpurun): bash tools/gpu_run.sh [tests|bench|all]
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == tests || $what == all ]]; then
  timeout 1200 python -m pytest tests -q -m gpu --maxfail=12 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
fi
if [[ $what == bench || $what == all ]]; then
  timeout 900 python bench.py --steps 3 --warmup 3 --cpu-timeout 200 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
