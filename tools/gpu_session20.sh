#!/bin/bash
# attention backward: one dQ MMA burst per pair of query blocks (B2PC_ATTN_RING=6) vs the default (4); parity first, then the
# full GPU suite and a short bench with the faster parity-green one
mkdir -p gpurun_out
bwd_ms() { grep -m1 "H=2 " $1 | sed -E 's/.*bwd impl2=([0-9.]+)ms.*/\1/'; }
for r in 6 4; do B2PC_ATTN_RING=$r PROBE_FAST=1 timeout 60 python tools/probe_attn.py time > gpurun_out/probe_attn_ring$r.log 2>&1; echo "ring $r:"; cut -c1-200 gpurun_out/probe_attn_ring$r.log | tail -3; done
B2PC_ATTN_RING=6 timeout 120 python -m pytest tests -q -m gpu -x -k "attn or attention or ptv3 or serialized or flash" 2>&1 | tail -5 > gpurun_out/pytest_ring6.log; tail -2 gpurun_out/pytest_ring6.log
T6=$(bwd_ms gpurun_out/probe_attn_ring6.log); T4=$(bwd_ms gpurun_out/probe_attn_ring4.log)
if grep -q " passed" gpurun_out/pytest_ring6.log && ! grep -q "failed\|error" gpurun_out/pytest_ring6.log && python -c "import sys; sys.exit(0 if float('$T6') < 0.985 * float('$T4') else 1)"; then
  export B2PC_ATTN_RING=6; echo "== ring 6 is parity-green and faster ($T6 < $T4 ms): used below"
else
  echo "== ring 6 rejected ($T6 vs $T4 ms): default ring 4 used below"
fi
echo "B2PC_ATTN_RING=${B2PC_ATTN_RING:-4}" > gpurun_out/ring_choice.txt
timeout 200 python -m pytest tests -q -m gpu --maxfail=10 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -1 gpurun_out/bench_short.err
