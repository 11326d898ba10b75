#!/bin/bash
# final-state validation on one B200: attention-backward ring A/B, full GPU suite, smoke, the gradient exchange at world size 1,
# the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for r in 4 3; do B2PC_ATTN_RING=$r PROBE_FAST=1 timeout 100 python tools/probe_attn.py time > gpurun_out/probe_attn_ring$r.log 2>&1; echo "ring $r:"; cut -c1-230 gpurun_out/probe_attn_ring$r.log | tail -3; done
timeout 420 python -m pytest tests -q -m gpu --maxfail=10 --durations=8 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -16 gpurun_out/pytest_gpu.log
if ! grep -q " passed" gpurun_out/pytest_gpu.log || grep -q " failed" gpurun_out/pytest_gpu.log; then
  echo "== failures: re-running the failed tests with B2PC_ATTN_RING=3"
  B2PC_ATTN_RING=3 timeout 300 python -m pytest tests -q -m gpu --maxfail=10 --lf 2>&1 | tail -15 > gpurun_out/pytest_gpu_ring3.log; tail -6 gpurun_out/pytest_gpu_ring3.log
fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1], 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value'] or 0), 'launches', d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],4), 'gref', (d.get('gpu_reference') or {}).get('ms_per_step'), d['config'].get('grad_exchange'), d['clocks'])" $1; }
timeout 150 python bench.py --steps 40 --warmup 5 --force-dist --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_forcedist_flat.json 2> gpurun_out/bench_forcedist_flat.err; tail -2 gpurun_out/bench_forcedist_flat.err; show gpurun_out/bench_forcedist_flat.json
B2PC_ATTN_RING=3 timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference > gpurun_out/bench_ring3.json 2> gpurun_out/bench_ring3.err; tail -1 gpurun_out/bench_ring3.err; show gpurun_out/bench_ring3.json
timeout 400 python bench.py --torch-profile gpurun_out/torch_profile_step.txt > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; show gpurun_out/bench_default.json
head -12 gpurun_out/torch_profile_step.txt | cut -c1-140
