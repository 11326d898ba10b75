#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_data.py tests/test_gpu_eval_ops.py -q 2>&1 | tail -30 > gpurun_out/pytest_data.log; tail -4 gpurun_out/pytest_data.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference --torch-profile gpurun_out/torch_profile_step.txt > gpurun_out/bench_default.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
head -70 gpurun_out/torch_profile_step.txt | cut -c1-170
