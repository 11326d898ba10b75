#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
SECONDS=0
timeout 1200 python -m pytest tests -q -m gpu --maxfail=12 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log; echo "pytest took $SECONDS s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
SECONDS=0
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_default.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_default.log | cut -c1-600; tail -2 gpurun_out/bench.err
echo "bench took $SECONDS s"
