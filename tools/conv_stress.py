"""Shape sweep of the warp-specialised sparse-conv kernels against the SIMT reference kernels (same C ABI), forward, backward-data
and weight gradient; prints one line per shape as it goes (run under `timeout`: a hang shows as the last line printed).
    python tools/conv_stress.py [quick]"""
import os
import sys
import time

import numpy as np
import torch

os.environ.setdefault("B2PC_CONV_WS", "1")    # sweep the warp-specialised forward kernel too (not the default)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcept_b200 import ops, synth  # noqa: E402

DEV = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))


def levels():
    b = synth.make_batch(2, seed=100)
    g, bb = b["grid_coord"], np.repeat(np.arange(2), np.diff(b["offset"], prepend=0))
    out = []
    for s in range(5):
        out.append((g, bb))
        key = np.concatenate([bb[:, None], g >> 1], 1)
        u = np.unique(key, axis=0)
        bb, g = u[:, 0], u[:, 1:]
    return out


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    lv = levels()
    # (level, cin, cout, ksize): PT-v3 CPE widths, SpUNet widths incl. the concatenated decoder inputs, stem
    shapes = [(0, 32, 32, 3), (0, 64, 64, 3), (1, 64, 64, 3), (2, 128, 128, 3), (3, 256, 256, 3), (4, 512, 512, 3), (0, 16, 32, 5),
              (0, 128, 96, 3), (1, 192, 128, 3), (2, 384, 256, 3), (3, 256, 256, 3), (0, 96, 96, 3), (1, 224, 96, 3), (2, 48, 32, 3)]
    if quick:
        shapes = shapes[:7]
    bad = 0
    for lvl, cin, cout, ks in shapes:
        g, bb = lv[lvl]
        n = len(g)
        idx = torch.from_numpy(np.concatenate([bb[:, None], g], 1).astype(np.int32)).to(DEV)
        pair = ops.rulebook_subm(idx, (g.max(0) + 96).tolist(), ks)
        kv = pair.shape[0]
        torch.manual_seed(cin + cout)
        feat = torch.randn(n, cin, device=DEV).bfloat16()
        w = (torch.randn(cout, kv, cin, device=DEV) / np.sqrt(cin * 11)).bfloat16()
        dout = torch.randn(n, cout, device=DEV).bfloat16()
        res = {}
        t0 = time.time()
        for impl in (1, 2):
            ops.set_impl(impl)
            stage = "fwd"
            try:
                fwd = ops._gather_gemm(feat, w, None, pair, n, cin, cout, kv, False, False)
                torch.cuda.synchronize()
                stage = "bwd_data"
                bwd = ops._gather_gemm(dout, w, None, pair, n, cout, cin, kv, True, True)
                torch.cuda.synchronize()
                stage = "bwd_weight"
                dw = torch.empty((cout, kv, cin), dtype=torch.float32, device=DEV)
                L = ops._lib.lib()
                ws = ops._ws(L.b2pc_spconv_bwd_weight_workspace_bytes(n, cin, cout, kv), DEV)
                ops._lib.check(L.b2pc_spconv_bwd_weight(ops._p(feat), ops._p(dout), ops._p(pair), pair.shape[1], n, n, cin, cout, kv, 2,
                                                        ops._p(dw), ops._p(ws), ws.numel(), impl, ops._stream()), "bwd_weight")
                torch.cuda.synchronize()
            except Exception as e:
                print(f"N={n} {cin}->{cout} kv={kv} impl={impl}: FAILED in {stage}: {str(e)[:120]}", flush=True)
                sys.exit(2)
            res[impl] = (fwd.float(), bwd.float(), dw)
        ops.set_impl(0)
        e = [rel(a, b) for a, b in zip(res[2], res[1])]
        ok = all(v < 3e-3 for v in e)
        bad += not ok
        print(f"N={n:7d} {cin:3d}->{cout:3d} kv={kv:3d}: fwd {e[0]:.1e} bwd_data {e[1]:.1e} dW {e[2]:.1e}  {'ok' if ok else 'MISMATCH'}  "
              f"({time.time() - t0:.1f}s)", flush=True)
    print("ALL OK" if not bad else f"{bad} MISMATCHES", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
