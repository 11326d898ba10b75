#!/usr/bin/env python
"""points/sec forward+backward of PT-v3m1 "base" (configs/scannet/semseg-pt-v3m1-0-base.py) on synthetic
ScanNet-scale scenes, sharded by whole scenes over N GPUs (DDP, NCCL gradient all-reduce only).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU restatement of the reference path, host cores

One step = zero_grad + forward (bf16 autocast) + loss + backward (+ DDP all-reduce) + AdamW step on a batch of
`--scenes-per-gpu` scenes per rank (weak scaling; BASELINE config 4: batch 16 over 8 GPUs = 2 scenes / GPU).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "points/sec fwd+bwd (PTv3-base, ScanNet-scale synth)"   # --workload spunet34 reports the same unit for SpUNet-34
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes-per-gpu", type=int, default=2)
    ap.add_argument("--voxels", type=int, default=120_000, help="voxels per synthetic scene")
    ap.add_argument("--cpu-voxels", type=int, default=120_000, help="scene size of the bounded CPU sample (one scene)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-timeout", type=int, default=240)
    ap.add_argument("--workload", default="ptv3_base", choices=["ptv3_base", "spunet34"],
                    help="ptv3_base = the headline metric (BASELINE config 4 shape); spunet34 = BASELINE config 3 (supplementary)")
    ap.add_argument("--fused-linear", action="store_true", help="fused bias-gradient Linear (pays off for GPU-bound batches)")
    ap.add_argument("--no-reorder", action="store_true", help="keep level-0 points in input order (no z-order memory layout)")
    ap.add_argument("--kernel-impl", type=int, default=None, help="0 auto, 1 SIMT kernels, 2 tcgen05 kernels")
    return ap.parse_args()


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def host_cores():
    """usable host cores: affinity mask, further limited by a cgroup cpu quota if one is set"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return max(1, n)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback")


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference path, all host threads, bounded sample
# ---------------------------------------------------------------------------------------------------------
def cpu_step_fn(voxels, seed=0):
    import numpy as np
    import torch
    from oracle import ptv3_cpu
    from pointcept_b200 import synth
    from pointcept_b200.ptv3 import PTv3Segmentor, ptv3_base_config

    cores = host_cores()
    torch.set_num_threads(cores)
    cfg = ptv3_base_config()
    torch.manual_seed(0)
    model = PTv3Segmentor(num_classes=20, backbone_out_channels=64, **cfg)   # parameters only; never executed on CPU
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    batch = synth.make_batch(1, seed=seed, target_voxels=voxels)
    n = int(batch["offset"][-1])

    def step():
        for v in sd.values():
            v.grad = None
        feat = ptv3_cpu.forward(bsd, dict(grid_coord=batch["grid_coord"], feat=batch["feat"], offset=batch["offset"]), cfg, bn_training=True)
        logits = torch.nn.functional.linear(feat, sd["seg_head.weight"], sd["seg_head.bias"])
        loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(batch["segment"]))
        loss.backward()
        return float(loss.detach())

    return step, n, cores


def run_reference(args):
    """--impl reference: the reference's own CPU path (its serialization / padding / dense attention math as restated
    and pinned in oracle/, spconv restated) timed on the host cores.  Only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    step, n, cores = cpu_step_fn(args.cpu_voxels)
    log(f"reference arm: {n} voxels on {cores} host cores")
    t0 = time.perf_counter()
    step()                                   # warm-up (also tells us how long a step takes)
    warm = time.perf_counter() - t0
    log(f"warm-up step {warm:.1f}s")
    steps = max(1, min(args.steps, 3, int(90.0 / max(warm, 1e-3))))   # keep the whole arm within a few minutes
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = n / dt
    sample = f"1 scene x {n} voxels, PTv3-base fwd+bwd fp32, torch CPU, {steps} timed step(s)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 1,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "PT-v3m1 base (configs/scannet/semseg-pt-v3m1-0-base.py), bounded CPU sample", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pointcept_b200 import _lib, ops, synth
    from pointcept_b200.ptv3 import PTv3Segmentor, ptv3_base_config

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 operators have no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.kernel_impl is not None:
        ops.set_impl(args.kernel_impl)

    if args.fused_linear:
        from pointcept_b200.ptv3 import FusedLinear
        FusedLinear.use_fused_bias_grad = True
    torch.manual_seed(0)
    if args.workload == "spunet34":
        from pointcept_b200.spunet import SpUNetBase

        class _SpUNetSeg(torch.nn.Module):
            """SpUNet-v1m1 stock widths (configs/scannet/semseg-spunet-v1m1-0-base.py:10-20) + CE loss"""

            def __init__(self):
                super().__init__()
                self.backbone = SpUNetBase(6, 20)

            def prepare(self, d):
                return d

            def forward(self, d):
                logits = self.backbone(d)
                return dict(seg_logits=logits, loss=torch.nn.functional.cross_entropy(logits.float(), d["segment"]))

        model = _SpUNetSeg().to(dev).train()
    else:
        model = PTv3Segmentor(num_classes=20, backbone_out_channels=64, spatial_reorder=not args.no_reorder,
                              **ptv3_base_config()).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False,
                                                        gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=0.05, fused=True)

    # synthetic batch of this rank (weak scaling: per-GPU work fixed), resident in pinned host memory
    hb = synth.make_batch(args.scenes_per_gpu, seed=100 + rank, target_voxels=args.voxels)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in hb.items()}
    offset_host = [int(v) for v in hb["offset"]]
    grid_max_host = [int(v) for v in hb["grid_coord"].max(0)]
    n_points = offset_host[-1]
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned.values())

    def to_device():
        d = {k: t.to(dev, non_blocking=True) for k, t in pinned.items()}
        d["offset_host"], d["grid_max_host"] = offset_host, grid_max_host   # host metadata the collate already has
        return d

    def step(d):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(d)
        out["loss"].backward()
        opt.step()
        return out["loss"]

    # Coordinate-only preparation (serialization, pooling index plans: the model's only host syncs) of the NEXT batch runs on
    # a side stream while the current batch trains, the way a data loader prefetches: its syncs then wait for the small
    # side-stream queue instead of the whole training backlog.  Work is done every step (nothing is cached across steps).
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def prepare_async(make_inputs):
        with torch.cuda.stream(side):
            d = make_inputs()
            point = model.prepare(d)
            point["segment"] = d["segment"]
            ev = torch.cuda.Event()
            ev.record(side)
        if hasattr(point, "record_stream"):
            point.record_stream(main)
        else:
            for v in point.values():
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(main)
        return point, ev

    def run_steps(n, make_inputs, on_loss=None):
        nxt = prepare_async(make_inputs)
        for i in range(n):
            point, ev = nxt
            if i + 1 < n:
                nxt = prepare_async(make_inputs)
            main.wait_event(ev)
            loss = step(point)
            if on_loss is not None:
                on_loss(i, loss)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = to_device()
    torch.cuda.synchronize()

    def resident_inputs():
        return dict(resident)

    log(f"rank {rank}: model on device, {n_points} points per step; warm-up")
    run_steps(max(args.warmup, 3), resident_inputs)
    sync_all()
    log("warm-up done; timed region 1")

    # ---- timed region 1: inputs resident in HBM, CUDA events, max over ranks ------------------------------
    launches0 = _lib.lib().b2pc_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    run_steps(args.steps, resident_inputs)
    e1.record()
    sync_all()
    clocks = sampler.stop() if sampler else None
    launches = _lib.lib().b2pc_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    pts = torch.tensor([float(n_points)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(pts, op=dist.ReduceOp.SUM)
    total_points = float(pts.item())
    value = total_points * args.steps / (ms_total * 1e-3)

    # ---- roofline leg: the same steps again with a CUDA-event pair around every launch of the hot operators ----
    # (kept out of region 1 so that the event bookkeeping does not tax the headline number)
    ops.profile_start()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_prof = min(args.steps, 3)
    sync_all()
    p0.record()
    run_steps(n_prof, resident_inputs)
    p1.record()
    sync_all()
    prof = ops.profile_stop()
    prof_ms_total = p0.elapsed_time(p1)
    log(f"timed region 1: {ms_total / args.steps:.1f} ms/step; timed region 2 (e2e)")
    # ---- timed region 2: end to end through the public API with HOST buffers ---------------------------------
    loss_pinned = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    sync_all()
    t0 = time.perf_counter()
    def read_loss(i, loss):                                    # device -> host read of the step's result
        loss_pinned[i:i + 1].copy_(loss.detach().float().reshape(1), non_blocking=True)

    run_steps(args.steps, to_device, read_loss)                # host -> device copy of every step's inputs inside prepare_async
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    loss_host = float(loss_pinned[-1])
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_points * args.steps / float(t.item())

    # ---- roofline of the dominant instrumented kernel -------------------------------------------------------
    pk = peaks()
    roof, shares = None, {}
    if prof:
        tot = {}
        for name, recs in prof.items():
            tot[name] = sum(a.elapsed_time(b) for a, b, _ in recs)
        shares = {k: round(v / prof_ms_total, 4) for k, v in tot.items()}
        top = max(tot, key=tot.get)
        recs = prof[top]
        if top.startswith("patch_attn"):
            flops = 0.0
            for _, _, (cu, H, D) in recs:
                lens = torch.diff(cu).double()
                flops += float((lens * lens).sum().item()) * H * D * (4.0 if top.endswith("fwd") else 10.0)
            ach = flops / (tot[top] * 1e-3) / 1e12
            roof = dict(kernel=top, bound="tensor", achieved=ach, peak=pk["tensor_sustained"], unit="TFLOP/s",
                        frac=ach / pk["tensor_sustained"], traffic=None, launches=len(recs), avg_ms=tot[top] / len(recs),
                        peak_source=pk["source"] + " (sustained cuBLAS bf16)",
                        note="D=16 attention is exp-pipe bound (64 MMA-flop per exp): see DESIGN.md")
        else:
            byts = 0.0
            for _, _, (pair, cin, cout, es) in recs:
                valid = float((pair >= 0).sum().item())
                n_out = pair.shape[1]
                byts += valid * cin * es + n_out * cout * es + pair.numel() * 4 + pair.shape[0] * cin * cout * es
            ach = byts / (tot[top] * 1e-3) / 1e9
            roof = dict(kernel=top, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], traffic=None,
                        launches=len(recs), avg_ms=tot[top] / len(recs), peak_source=pk["source"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        log("cpu_baseline: running the oracle port on the host cores (bounded subprocess)")
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1",
                                "--cpu-voxels", str(args.cpu_voxels)], capture_output=True, text=True, timeout=args.cpu_timeout,
                               env=dict(os.environ, RANK="0", WORLD_SIZE="1", CUDA_VISIBLE_DEVICES=""))
            js = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            cpu = json.loads(js[-1])["cpu_baseline"] if js else dict(value=None, unit=UNIT, error=p.stderr[-300:])
        except subprocess.TimeoutExpired:
            cpu = dict(value=None, unit=UNIT, cores=host_cores(), kind="port",
                       sample=f"1 scene x {args.cpu_voxels} voxels", error=f"did not finish within {args.cpu_timeout}s")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("PT-v3m1 base (configs/scannet/semseg-pt-v3m1-0-base.py) fwd+bwd+AdamW, BASELINE config 4 shape: "
                                if args.workload == "ptv3_base" else
                                "SpUNet-v1m1 34 (configs/scannet/semseg-spunet-v1m1-0-base.py) fwd+bwd+AdamW, BASELINE config 3 shape: ")
                               + f"{args.scenes_per_gpu} synthetic ScanNet-scale scenes per GPU",
                   "scenes_per_gpu": args.scenes_per_gpu, "points_per_gpu": n_points, "global_points": int(total_points),
                   "params_M": round(n_params / 1e6, 2), "patch_size": 1024, "orders": 4, "parallelism": f"dp{world}",
                   "l2": "no explicit flush: one step streams several GB of activations, far beyond the 126 MB L2",
                   "kernel_impl": ops.get_impl(), "spatial_reorder": not args.no_reorder, "loss_last": loss_host},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernel_time_share": shares,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
