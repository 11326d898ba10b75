"""AdamW over the whole parameter list in one kernel launch ("fused AdamW behind the all-reduce").

Same update rule, hyper-parameters and per-parameter state (``exp_avg`` / ``exp_avg_sq`` / ``step``) as ``torch.optim.AdamW`` --
the optimizer the reference builds (pointcept/utils/optimizer.py:12-58, configs ``optimizer = dict(type="AdamW", ...)``) -- but
the moments live in two flat fp32 buffers and one launch of ``b2pc_multi_adamw`` updates every tensor: torch's fused path still
issues one multi-tensor launch per ~30 tensors (~80 launches for PT-v3-base's 486 tensors).  Per step the host only refreshes a
table of gradient pointers and bias corrections (one small pinned -> device copy).
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._plans = {}

    def _plan(self, gi, group):
        params = [p for p in group["params"] if p.requires_grad]
        key = tuple(p.data_ptr() for p in params)
        pl = self._plans.get(gi)
        if pl is None or pl["key"] != key:
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdamW takes contiguous fp32 CUDA parameters")
            dev = params[0].device
            sizes = [(p.numel() + 3) // 4 * 4 for p in params]
            m = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            v = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            base = np.zeros((len(params), 5), dtype=np.int64)       # p, m, v, count, blocks
            steps = np.zeros(len(params), dtype=np.int64)
            off = 0
            for i, (p, sz) in enumerate(zip(params, sizes)):
                st = self.state[p]
                mi, vi = m[off:off + p.numel()].view_as(p), v[off:off + p.numel()].view_as(p)
                if "exp_avg" in st:            # state loaded from a checkpoint / carried over from an earlier plan
                    mi.copy_(st["exp_avg"])
                    vi.copy_(st["exp_avg_sq"])
                    steps[i] = int(st.get("step", 0))
                st["exp_avg"], st["exp_avg_sq"] = mi, vi
                base[i] = (p.data_ptr(), mi.data_ptr(), vi.data_ptr(), p.numel(), (p.numel() + 2047) // 2048)
                off += sz
            pinned = torch.zeros((len(params), 7), dtype=torch.int64).pin_memory()
            pl = dict(key=key, params=params, m=m, v=v, base=base, steps=steps, pinned=pinned,
                      dev_table=torch.empty((len(params), 7), dtype=torch.int64, device=dev))
            self._plans[gi] = pl
        return pl

    def state_dict(self):
        for pl in self._plans.values():          # per-parameter step counts live in the plan; publish them torch-style on demand
            for p, t in zip(pl["params"], pl["steps"]):
                self.state[p]["step"] = torch.tensor(float(t))
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        wrote = False
        for gi, group in enumerate(self.param_groups):
            pl = self._plan(gi, group)
            params, base = pl["params"], pl["base"]
            live = np.fromiter((i for i, p in enumerate(params) if p.grad is not None), dtype=np.int64)
            if live.size == 0:
                continue
            gptr = np.empty(live.size, dtype=np.int64)
            for n, i in enumerate(live):
                g = params[i].grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    raise RuntimeError("FusedAdamW takes contiguous fp32 gradients")
                gptr[n] = g.data_ptr()
            pl["steps"][live] += 1
            b1, b2 = group["betas"]
            t = pl["steps"][live].astype(np.float64)
            tab = pl["pinned"].numpy()[:live.size]
            tab[:, 0] = base[live, 0]
            tab[:, 1] = gptr
            tab[:, 2] = base[live, 1]
            tab[:, 3] = base[live, 2]
            tab[:, 4] = base[live, 3]
            blocks = base[live, 4]
            tab[:, 5] = np.cumsum(blocks) - blocks
            bc = np.stack([1.0 - b1 ** t, np.sqrt(1.0 - b2 ** t)], 1).astype(np.float32)      # two floats packed into the 7th word
            tab[:, 6] = bc.view(np.int64)[:, 0]
            pl["dev_table"][:live.size].copy_(pl["pinned"][:live.size], non_blocking=True)
            _lib.check(L.b2pc_multi_adamw(ctypes.c_void_p(pl["dev_table"].data_ptr()), int(live.size), int(blocks.sum()), float(group["lr"]),
                                          float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), float(grad_scale),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "multi_adamw")
            wrote = True
        if wrote:
            # the kernel wrote the parameters in place behind autograd's back (no version-counter bump): tell the half-precision
            # shadows (ops.HalfShadows) that every parameter may have changed
            ops.bump_param_epoch()
        return loss
