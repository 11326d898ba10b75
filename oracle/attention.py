"""torch-fp32 restatement of PTv3 patch attention (TEST ORACLE).

The math of flash_attn.flash_attn_varlen_qkvpacked_func as the reference calls it
(point_transformer_v3m1_base.py:208-214): for every sequence s in cu_seqlens and head h,
out = softmax(scale * Q K^T) V, no mask, no dropout.  Equals the reference's non-flash
branch (":190-206") when every sequence has length K; pinned against
tests/golden/attention_dense.npz generated from that branch.
"""
import torch


def varlen_attention(qkv, cu_seqlens, scale=None, return_lse=False):
    """qkv [T,3,H,D] (any float dtype; math in fp32), cu_seqlens int [n+1] -> out [T,H,D] fp32, lse [H,T]."""
    T, three, H, D = qkv.shape
    assert three == 3
    scale = D ** -0.5 if scale is None else scale
    q, k, v = qkv.float().unbind(1)
    out = torch.empty(T, H, D, dtype=torch.float32)
    lse = torch.empty(H, T, dtype=torch.float32)
    cu = [int(c) for c in cu_seqlens]
    for a, b in zip(cu[:-1], cu[1:]):
        if b == a:
            continue
        s = torch.einsum("qhd,khd->hqk", q[a:b] * scale, k[a:b])
        lse[:, a:b] = torch.logsumexp(s, dim=-1)
        out[a:b] = torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), v[a:b])
    return (out, lse) if return_lse else out


def varlen_attention_grads(qkv, cu_seqlens, dout, scale=None):
    """dqkv [T,3,H,D] fp32 by autograd through the dense math."""
    x = qkv.detach().float().requires_grad_(True)
    out = varlen_attention(x, cu_seqlens, scale)
    out.backward(dout.float())
    return x.grad
