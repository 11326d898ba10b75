#!/bin/bash
# gradient exchange A/B on two B200s of one box: FlatGradReducer vs torch DDP, same build, same box, plus the N = 1 line of that box
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1], 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value'] or 0), d['config'].get('grad_exchange'), d['clocks'])" $1; }
COMMON="--steps 40 --warmup 5 --no-cpu-baseline --no-supplementary --no-gpu-reference"
timeout 100 python bench.py --gpus 1 $COMMON > gpurun_out/s18_n1.json 2> gpurun_out/s18_n1.err; show gpurun_out/s18_n1.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 $COMMON --grad-exchange flat > gpurun_out/s18_n2_flat.json 2> gpurun_out/s18_n2_flat.err; show gpurun_out/s18_n2_flat.json; grep -m3 -i "NVLS\|nranks" gpurun_out/s18_n2_flat.err | cut -c1-200
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 $COMMON --grad-exchange ddp > gpurun_out/s18_n2_ddp.json 2> gpurun_out/s18_n2_ddp.err; show gpurun_out/s18_n2_ddp.json
tail -3 gpurun_out/s18_n2_flat.err | cut -c1-300
