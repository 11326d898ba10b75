"""Variant / evaluation-time operators through the C ABI: kNN query, fragment voting, PointROPE, and patch attention
at LitePT's head_dim 18.  Tolerances are stated at each assert."""
import os

import numpy as np
import pytest
import torch

from oracle import attention as oattn
from oracle import eval_ops as oev
from pointcept_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("k", [1, 3, 8, 16, 40])
def test_knn_query_matches_oracle(k):
    rng = np.random.default_rng(k)
    sizes, qsizes = [700, 1500, 33], [400, 900, 50]
    xyz = rng.random((sum(sizes), 3)).astype(np.float32) * 3
    new = rng.random((sum(qsizes), 3)).astype(np.float32) * 3
    off, qoff = np.cumsum(sizes), np.cumsum(qsizes)
    idx, dist = ops.knn_query(k, torch.from_numpy(xyz).to(DEV), torch.from_numpy(off).to(DEV), torch.from_numpy(new).to(DEV),
                              torch.from_numpy(qoff).to(DEV))
    widx, wdist = oev.knn_query(k, xyz, off, new, qoff)
    assert idx.dtype == torch.int32 and idx.shape == (sum(qsizes), k)
    # distances: fp32 rounding of the squared distance only (FMA contraction differs between numpy and the kernel): 1e-6 relative
    assert np.allclose(dist.cpu().numpy(), wdist, rtol=1e-5, atol=1e-7)
    same = idx.cpu().numpy() == widx
    assert same.mean() > 0.999          # an index may differ only where two candidates are equidistant to the last bit
    far = ~same
    assert np.allclose(dist.cpu().numpy()[far], wdist[far], rtol=1e-5, atol=1e-7)
    if k == 40:                          # third scene has 33 < 40 points: placeholders exactly as the reference's
        assert (idx[-50:, 33:] == -1).all() and torch.allclose(dist[-50:, 33:], torch.full((50, 7), 1e5, device=DEV))


def test_knn_self_query_and_scene_isolation():
    rng = np.random.default_rng(0)
    xyz = torch.from_numpy(rng.random((5000, 3)).astype(np.float32)).to(DEV)
    off = torch.tensor([2000, 5000], device=DEV)
    idx, dist = ops.knn_query(4, xyz, off)
    assert (idx[:, 0].long() == torch.arange(5000, device=DEV)).all() and (dist[:, 0] == 0).all()
    assert (idx[:2000] < 2000).all() and (idx[2000:] >= 2000).all()
    assert (dist[:, 1:] >= dist[:, :-1]).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_vote_accumulate(dtype):
    gen = torch.Generator().manual_seed(0)
    n_all, c = 5000, 20
    pred = torch.rand(n_all, c, generator=gen)
    index = torch.randperm(n_all, generator=gen)[:3000]
    logits = (torch.randn(3000, c, generator=gen) * 3).to(dtype)
    got = ops.vote_accumulate(pred.clone().to(DEV), index.to(DEV), logits.to(DEV)).cpu()
    want = oev.vote(pred, index, logits)
    assert float((got - want).abs().max()) <= 2e-6        # fp32 softmax of the same (possibly half-precision) logits
    # 200-class case (ScanNet200) and repeated indices
    logits = torch.randn(64, 200, generator=gen)
    index = torch.randint(0, 10, (64,), generator=gen)
    got = ops.vote_accumulate(torch.zeros(10, 200, device=DEV), index.to(DEV), logits.to(DEV)).cpu()
    assert float((got - oev.vote(torch.zeros(10, 200), index, logits)).abs().max()) <= 1e-5


def test_point_rope_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "point_rope.npz"))
    from pointcept_b200 import pointrope
    for name in ("d18", "d48"):
        tok = torch.from_numpy(g[name + "_tokens"])[None].contiguous().to(DEV)           # [1, N, H, D]
        pos = torch.from_numpy(g[name + "_pos"])[None].contiguous().to(DEV)
        want = torch.from_numpy(g[name + "_out"])
        out = pointrope.pointrope(tok.clone(), pos, float(g[name + "_base"]), 1.0)
        # positions up to 700 -> angles up to 700 rad: cosf / sinf of the device vs torch's CPU kernels, a few ulp of the angle
        assert float((out[0].cpu() - want).abs().max()) <= 2e-4, name
        back = pointrope.pointrope(out.clone(), pos, float(g[name + "_base"]), -1.0)
        assert float((back - tok).abs().max()) <= 2e-4
        # half precision tokens: fp32 math, one rounding on the way out
        out16 = pointrope.pointrope(tok.half(), pos, float(g[name + "_base"]), 1.0)
        assert float((out16[0].float().cpu() - want).abs().max()) <= 4e-3 * float(want.abs().max())


def test_rope_qkv_fused_forward_backward():
    gen = torch.Generator().manual_seed(3)
    t, h, d = 500, 4, 18
    qkv = torch.randn(t, 3, h, d, generator=gen)
    pos = torch.randint(0, 300, (t, 3), generator=gen)
    x = qkv.to(DEV).requires_grad_(True)
    y = ops.rope_qkv(x, pos.to(DEV), 100.0, 1.0)
    wq, wk = oev.point_rope(qkv[:, 0], pos, 100.0), oev.point_rope(qkv[:, 1], pos, 100.0)
    assert float((y[:, 0].detach().cpu() - wq).abs().max()) <= 1e-4 and float((y[:, 1].detach().cpu() - wk).abs().max()) <= 1e-4
    assert torch.equal(y[:, 2].detach().cpu(), qkv[:, 2])
    g = torch.randn(t, 3, h, d, generator=gen)
    y.backward(g.to(DEV))
    ref = qkv.clone().requires_grad_(True)
    (torch.stack([oev.point_rope(ref[:, 0], pos, 100.0), oev.point_rope(ref[:, 1], pos, 100.0), ref[:, 2]], 1) * g).sum().backward()
    assert float((x.grad.cpu() - ref.grad).abs().max()) <= 1e-4


@pytest.mark.parametrize("d", [18, 24, 48])
def test_patch_attention_litept_head_dims(d):
    """LitePT feeds fp16 packed qkv with head_dim 18 to flash_attn_varlen_qkvpacked_func (litept_v1.py:252-259)."""
    gen = torch.Generator().manual_seed(d)
    lens = [200, 77, 128]
    t, h = sum(lens), 3
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
    qkv = torch.randn(t, 3, h, d, generator=gen)
    from pointcept_b200.flash_attn_interface import flash_attn_varlen_qkvpacked_func
    x = qkv.half().to(DEV).requires_grad_(True)
    out = flash_attn_varlen_qkvpacked_func(x, cu.to(DEV), max(lens), softmax_scale=d ** -0.5)
    want = oattn.varlen_attention(qkv.half(), cu, d ** -0.5)
    assert float((out.detach().float().cpu() - want).abs().max()) <= 2e-3          # fp16 output rounding
    g = torch.randn(t, h, d, generator=gen)
    out.backward(g.half().to(DEV))
    wgrad = oattn.varlen_attention_grads(qkv.half(), cu, g.half(), d ** -0.5)
    assert float((x.grad.float().cpu() - wgrad).abs().max()) <= 4e-3 * max(1.0, float(wgrad.abs().max()))
