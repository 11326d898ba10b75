"""Quick A/B of the TMA (gather4) conv path against the SIMT kernels of the same library, small shapes (run under gpurun with a timeout)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcept_b200 import ops  # noqa: E402

dev = "cuda"
rng = np.random.default_rng(0)
worst = 0.0
for n_target, c in ((2000, 32), (5000, 64), (1500, 128), (700, 256), (300, 512), (129, 64)):
    g = np.unique(rng.integers(0, 24, (n_target * 2, 3)), axis=0)[:n_target]
    idx = torch.from_numpy(np.concatenate([np.zeros((len(g), 1), np.int64), g], 1).astype(np.int32)).to(dev)
    pair = ops.rulebook_subm(idx, [120, 120, 120], 3)
    n, kv = len(g), pair.shape[0]
    feat = torch.randn(n, c, device=dev).bfloat16()
    w = (torch.randn(c, kv, c, device=dev) * 0.05).bfloat16()
    b = torch.randn(c, device=dev).bfloat16()
    res = {}
    for impl in (1, 2):
        ops.set_impl(impl)
        res[impl] = (ops._gather_gemm(feat, w, b, pair, n, c, c, kv, False, False).float(),
                     ops._gather_gemm(feat, w, None, pair, n, c, c, kv, True, True).float())
    torch.cuda.synchronize()
    for name, a, r in (("fwd", res[2][0], res[1][0]), ("bwd_data", res[2][1], res[1][1])):
        rel = float((a - r).norm() / r.norm())
        worst = max(worst, rel)
        print(f"n={n:5d} c={c:3d} {name}: rel {rel:.2e}", flush=True)
print("WORST", worst, "OK" if worst < 2e-2 else "MISMATCH")
