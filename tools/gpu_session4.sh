#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/conv_stress.py > gpurun_out/conv_stress.log 2>&1; rc=$?; echo "conv stress rc=$rc"; tail -16 gpurun_out/conv_stress.log | cut -c1-200
if [ $rc -ne 0 ]; then export B2PC_CONV_V1=1; echo "FALLING BACK TO V1 WGRAD FOR THE REST OF THIS SESSION"; fi
IMPLS=2 timeout 200 python tools/probe_conv.py > gpurun_out/probe_conv_default.log 2>&1; tail -7 gpurun_out/probe_conv_default.log
timeout 900 python bench.py --steps 50 --warmup 5 --cpu-timeout 150 --gpu-reference-steps 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
for f in test_gpu_fused test_gpu_ops test_gpu_scale_parity test_gpu_model; do
  timeout 500 python -m pytest tests/$f.py -q -rP 2>&1 | tail -60 > gpurun_out/pytest_$f.log; echo "$f: $(tail -1 gpurun_out/pytest_$f.log)"
done
grep -E "largest|zero ref|autocast torch|GradScaler" gpurun_out/pytest_test_gpu_model.log | cut -c1-500
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
# ncu --set full of the two warp-specialised conv kernels (C=32, N=240k) for offline analysis
B2PC_CONV_WS=1 ONLY=0 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ws_kernel -s 4 -c 1 -f -o gpurun_out/r02_conv_ws_c32 \
   python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log | cut -c1-200
ONLY=0 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_ws_kernel -s 2 -c 1 -f -o gpurun_out/r02_wgrad_ws_c32 \
   python tools/probe_conv.py > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log | cut -c1-200
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2600 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
