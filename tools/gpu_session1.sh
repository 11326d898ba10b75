#!/bin/bash
# round-2 GPU session 1: new warp-specialised conv kernels (parity first), new parity tests, ncu captures, B2 comparator.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
# (0) quick parity of the new conv kernels; if they fail or hang, the rest of the session runs on the round-1 kernels
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "conv or strided" > gpurun_out/pytest_conv_ws.log 2>&1
rc=$?; echo "conv_ws quick parity rc=$rc"; tail -3 gpurun_out/pytest_conv_ws.log
if [ $rc -ne 0 ]; then export B2PC_CONV_V1=1; echo "FALLING BACK TO V1 CONV KERNELS FOR THIS SESSION"; fi
# (1) can a spconv wheel be had on the box?  (no network: expected to fail; BASELINE.md B2 asks for the attempt to be recorded)
{ echo "== pip install spconv-cu124"; timeout 60 python -m pip install spconv-cu124 2>&1 | tail -4;
  echo "== pip install spconv-cu126"; timeout 60 python -m pip install spconv-cu126 2>&1 | tail -4;
  echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "spconv|cumm" || echo "no spconv/cumm wheel in /opt/wheelhouse";
  echo "== import"; python -c "import spconv; print('spconv', spconv.__version__)" 2>&1 | tail -1; } > gpurun_out/r02_spconv_install_attempt.txt 2>&1
# (2) per-op timings: new vs round-1 conv kernels, attention
IMPLS=2 timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv_ws.log 2>&1; tail -7 gpurun_out/probe_conv_ws.log
B2PC_CONV_V1=1 IMPLS=2 timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv_v1.log 2>&1; tail -7 gpurun_out/probe_conv_v1.log
timeout 300 python tools/probe_attn.py time > gpurun_out/probe_attn_time.log 2>&1; tail -4 gpurun_out/probe_attn_time.log
# (3) full parity suite
timeout 2400 python -m pytest tests -q -m gpu --maxfail=40 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
# (4) ncu --set full: attention backward (round-1 kernel, before its rewrite), new conv + wgrad kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_umma -s 3 -c 1 -f -o gpurun_out/r02_attn_bwd_before \
   python tools/probe_attn.py time > gpurun_out/ncu_attn_bwd.log 2>&1
ONLY=0 IMPLS=2 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_ws_kernel|wgrad_ws_kernel|bwd_weight_umma|gather_gemm_umma" -s 6 -c 3 -f -o gpurun_out/r02_conv_c32 \
   python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1
# (5) B2 comparator, whole model
timeout 900 python tools/time_reference.py --scenes 2 --steps 5 > gpurun_out/r02_time_reference.json 2> gpurun_out/time_reference.err
tail -c 600 gpurun_out/r02_time_reference.json; tail -3 gpurun_out/time_reference.err
# (6) bench
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-timeout 200 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
