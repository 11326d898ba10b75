#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/conv_stress.py > gpurun_out/conv_stress.log 2>&1; rc=$?; echo "conv stress rc=$rc"; tail -16 gpurun_out/conv_stress.log
IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_ws.log 2>&1; tail -7 gpurun_out/probe_conv_ws.log
if [ $rc -ne 0 ]; then export B2PC_CONV_V1=1; echo "FALLING BACK TO V1 CONV KERNELS FOR THE REST OF THIS SESSION"; fi
timeout 300 python -m pytest tests/test_gpu_model.py -q -rP -k "backward_all or autocast" 2>&1 | tail -60 > gpurun_out/pytest_model3.log; grep -E "largest|zero ref|autocast torch|passed|failed|Error" gpurun_out/pytest_model3.log | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_fused.py -q 2>&1 | tail -40 > gpurun_out/pytest_fused.log; tail -6 gpurun_out/pytest_fused.log
timeout 700 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --gpu-reference-steps 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
B2PC_BLOCK_FUSED=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_noblockfuse.log 2>&1; tail -c 700 gpurun_out/bench_noblockfuse.log | cut -c1-400
B2PC_ATTN_FUSED=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_noattnfuse.log 2>&1; tail -c 700 gpurun_out/bench_noattnfuse.log | cut -c1-400
B2PC_CONV_V1=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_convv1.log 2>&1; tail -c 700 gpurun_out/bench_convv1.log | cut -c1-400
for f in test_gpu_ops test_gpu_scale_parity test_gpu_model; do
  timeout 500 python -m pytest tests/$f.py -q 2>&1 | tail -30 > gpurun_out/pytest_$f.log; echo "$f: $(tail -1 gpurun_out/pytest_$f.log)"
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 2400 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
