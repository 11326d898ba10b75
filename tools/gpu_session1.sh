#!/bin/bash
# round-2 GPU session 1: baseline evidence for the kernels about to be rewritten + new parity tests + the B2 comparator.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
# (1) can a spconv wheel be had on the box?  (no network: expected to fail; BASELINE.md B2 asks for the attempt to be recorded)
{ echo "== pip install spconv-cu124 (no index reachable?)"; timeout 60 python -m pip install spconv-cu124 2>&1 | tail -4;
  echo "== pip install spconv-cu126"; timeout 60 python -m pip install spconv-cu126 2>&1 | tail -4;
  echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "spconv|cumm" || echo "no spconv/cumm wheel in /opt/wheelhouse";
  echo "== import"; python -c "import spconv; print('spconv', spconv.__version__)" 2>&1 | tail -1; } > gpurun_out/r02_spconv_install_attempt.txt 2>&1
# (2) parity
timeout 2400 python -m pytest tests -q -m gpu --maxfail=40 -x --co -q > /dev/null 2>&1
timeout 2400 python -m pytest tests -q -m gpu --maxfail=40 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
# (3) ncu --set full of the two kernels VERDICT r01 asks for (before the rewrite)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_umma -s 3 -c 1 -f -o gpurun_out/r02_attn_bwd_before \
   python tools/probe_attn.py time > gpurun_out/ncu_attn_bwd.log 2>&1
ONLY=0,5 IMPLS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_weight_umma -s 3 -c 1 -f -o gpurun_out/r02_wgrad_before \
   python tools/probe_conv.py > gpurun_out/ncu_wgrad.log 2>&1
# (4) B2 comparator, whole model
timeout 900 python tools/time_reference.py --scenes 2 --steps 5 > gpurun_out/r02_time_reference.json 2> gpurun_out/time_reference.err
tail -c 600 gpurun_out/r02_time_reference.json; tail -3 gpurun_out/time_reference.err
# (5) per-op timings (current kernels)
timeout 300 python tools/probe_attn.py time > gpurun_out/probe_attn_time.log 2>&1; tail -4 gpurun_out/probe_attn_time.log
IMPLS=2 timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv.log 2>&1; tail -7 gpurun_out/probe_conv.log
