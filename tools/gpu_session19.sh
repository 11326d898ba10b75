#!/bin/bash
# attention backward: software-pipelined MMA order (B2PC_ATTN_RING=5) vs the 4-deep ring (default); the faster parity-green one is
# exported for the rest: full GPU suite, ncu --set full of the kernel, the default bench line
mkdir -p gpurun_out
bwd_ms() { grep -m1 "H=2 " $1 | sed -E 's/.*bwd impl2=([0-9.]+)ms.*/\1/'; }
for r in 5 4; do B2PC_ATTN_RING=$r PROBE_FAST=1 timeout 100 python tools/probe_attn.py time > gpurun_out/probe_attn_ring$r.log 2>&1; echo "ring $r:"; cut -c1-200 gpurun_out/probe_attn_ring$r.log | tail -3; done
B2PC_ATTN_RING=5 timeout 240 python -m pytest tests -q -m gpu -x -k "attn or attention or ptv3 or serialized or flash" 2>&1 | tail -5 > gpurun_out/pytest_ring5.log; tail -2 gpurun_out/pytest_ring5.log
T5=$(bwd_ms gpurun_out/probe_attn_ring5.log); T4=$(bwd_ms gpurun_out/probe_attn_ring4.log)
if grep -q " passed" gpurun_out/pytest_ring5.log && ! grep -q "failed\|error" gpurun_out/pytest_ring5.log && python -c "import sys; sys.exit(0 if float('$T5') < float('$T4') else 1)"; then
  export B2PC_ATTN_RING=5; echo "== ring 5 is parity-green and faster ($T5 < $T4 ms): used below"
else
  echo "== ring 5 rejected ($T5 vs $T4 ms): default ring 4 used below"
fi
echo "B2PC_ATTN_RING=${B2PC_ATTN_RING:-4}" > gpurun_out/ring_choice.txt
timeout 300 python -m pytest tests -q -m gpu --maxfail=10 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
PROBE_FAST=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_umma -s 3 -c 1 -f -o gpurun_out/r02_attn_bwd_final python tools/probe_attn.py time > gpurun_out/ncu_attn_bwd.log 2>&1; tail -2 gpurun_out/ncu_attn_bwd.log | cut -c1-200
timeout 400 python bench.py --torch-profile gpurun_out/torch_profile_step.txt > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{\"metric\"')][-1]);print('default', 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value'] or 0), 'launches', d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],4), 'gref', (d.get('gpu_reference') or {}).get('ms_per_step'), d['clocks'])"
head -6 gpurun_out/torch_profile_step.txt | cut -c1-140
