#!/bin/bash
# round-2 GPU session 2: fused glue / serialized attention / pooling plan parity, new bench line, launch list, A/B of the fusions.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fused.py -q -x > gpurun_out/pytest_fused.log 2>&1; echo "fused tests rc=$?"; tail -5 gpurun_out/pytest_fused.log
timeout 2400 python -m pytest tests -q -m gpu --maxfail=40 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
# headline bench (everything on) + A/B of the two fusions (short runs, no extras)
timeout 1200 python bench.py --steps 50 --warmup 5 --cpu-timeout 200 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
B2PC_BLOCK_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_noblockfuse.log 2>&1; tail -c 600 gpurun_out/bench_noblockfuse.log
B2PC_ATTN_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_noattnfuse.log 2>&1; tail -c 600 gpurun_out/bench_noattnfuse.log
B2PC_CONV_V1=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_convv1.log 2>&1; tail -c 600 gpurun_out/bench_convv1.log
# launch list of one step (ncu, cold-cache serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
