"""GPU bring-up / timing probe for the tcgen05 attention kernels (run under gpurun)."""
import itertools
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcept_b200 import ops  # noqa: E402

DEV = "cuda"


def ref_attn(qkv, lens, scale):
    outs, lses = [], []
    s = 0
    for n in lens:
        q, k, v = qkv[s:s + n].float().unbind(1)
        a = torch.einsum("qhd,khd->hqk", q * scale, k)
        lses.append(torch.logsumexp(a, -1))
        outs.append(torch.einsum("hqk,khd->qhd", torch.softmax(a, -1), v))
        s += n
    return torch.cat(outs), torch.cat(lses, 1)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def correctness(bn):
    torch.manual_seed(0)
    H = 2
    for lens in ([128], [1024], [1024, 700, 64, 129, 1], [2048, 100]):
        T = sum(lens)
        cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=DEV)
        for dt in (torch.bfloat16, torch.float16):
            qkv = (torch.randn(T, 3, H, 16, device=DEV) * 1.5).to(dt)
            ro, rl = ref_attn(qkv, lens, 0.25)
            ops.set_impl(2)
            out, lse = ops.patch_attention(qkv, cu, max(lens), 0.25, return_lse=True)
            torch.cuda.synchronize()
            print(f"BN={bn} lens={lens} {dt}: out rel {rel(out.float(), ro):.3e}  lse maxabs {float((lse - rl).abs().max()):.3e}", flush=True)


def desc_scan(bn):
    torch.manual_seed(0)
    lens, H = [256], 1
    cu = torch.tensor([0, 256], dtype=torch.int32, device=DEV)
    qkv = (torch.randn(256, 3, H, 16, device=DEV)).to(torch.bfloat16)
    ro, _ = ref_attn(qkv, lens, 0.25)
    plane_q, plane_kv = 128 * 16, bn * 16
    ops.set_impl(2)
    for qs, ks, vs in itertools.product((0, 1), repeat=3):
        q = (plane_q, 128) if not qs else (128, plane_q)
        k = (plane_kv, 128) if not ks else (128, plane_kv)
        v = (128, plane_kv) if not vs else (plane_kv, 128)
        os.environ["B2PC_ATTN_DESC"] = f"{q[0]},{q[1]},{k[0]},{k[1]},{v[0]},{v[1]}"
        out = ops.patch_attention(qkv, cu, 256, 0.25)
        torch.cuda.synchronize()
        print(f"desc swap q={qs} k={ks} v={vs}: rel {rel(out.float(), ro):.3e}", flush=True)
    del os.environ["B2PC_ATTN_DESC"]


def timing():
    torch.manual_seed(0)
    for H, nseq in ((2, 236), (4, 236), (8, 16)):
        T = nseq * 1024
        cu = torch.arange(0, T + 1, 1024, dtype=torch.int32, device=DEV)
        qkv = torch.randn(T, 3, H, 16, device=DEV, dtype=torch.bfloat16, requires_grad=True)
        dout = torch.randn(T, H, 16, device=DEV, dtype=torch.bfloat16)

        def bench(fn, n=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        res = {}
        for impl in ((2,) if os.environ.get("PROBE_FAST") else (1, 2)):
            ops.set_impl(impl)
            res[f"fwd impl{impl}"] = bench(lambda: ops.patch_attention(qkv.detach(), cu, 1024, 0.25))
            out = ops.patch_attention(qkv, cu, 1024, 0.25)
            res[f"bwd impl{impl}"] = bench(lambda: out.backward(dout, retain_graph=True))
        try:
            import flash_attn
            res["fwd FA2"] = bench(lambda: flash_attn.flash_attn_varlen_qkvpacked_func(qkv.detach(), cu, 1024, softmax_scale=0.25))
            o2 = flash_attn.flash_attn_varlen_qkvpacked_func(qkv, cu, 1024, softmax_scale=0.25)
            res["bwd FA2"] = bench(lambda: o2.backward(dout, retain_graph=True))
        except Exception as e:
            res["FA2"] = str(e)[:80]
        exps = T * 1024 * H
        print(f"H={H} T={T}: " + "  ".join(f"{k}={v:.3f}ms" if isinstance(v, float) else f"{k}={v}" for k, v in res.items())
              + f"  | fwd exp-rate impl2 = {exps / res['fwd impl2'] / 1e9:.2f} Gexp/ms-> {exps / (res['fwd impl2'] * 1e-3) / 1e12:.2f} Texp/s", flush=True)


if __name__ == "__main__":
    bn = int(os.environ.get("B2PC_ATTN_BN", "128"))
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("scan", "all"):
        desc_scan(bn)
    if what in ("check", "all"):
        correctness(bn)
    if what in ("time", "all"):
        timing()
