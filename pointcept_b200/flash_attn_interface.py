"""``flash_attn`` drop-in for the one entry point Pointcept calls
(point_transformer_v3/point_transformer_v3m1_base.py:208-214; also m2/m3 and LitePT):
``flash_attn_varlen_qkvpacked_func`` with flash-attn 2.8.3's signature."""
from . import ops


def flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, causal=False,
                                     window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False,
                                     return_attn_probs=False):
    """qkv [T,3,H,D] fp16/bf16, cu_seqlens int32 [n+1] -> out [T,H,D].  Non-causal dense patch attention."""
    if dropout_p and dropout_p > 0.0:
        raise NotImplementedError("pointcept_b200 patch attention: dropout_p > 0 is not supported "
                                  "(every PT-v3 config sets attn_drop=0.0)")
    if causal or tuple(window_size) != (-1, -1) or softcap != 0.0 or alibi_slopes is not None:
        raise NotImplementedError("pointcept_b200 patch attention: causal/window/softcap/alibi are not supported")
    if return_attn_probs:
        out, lse = ops.patch_attention(qkv, cu_seqlens, max_seqlen, softmax_scale, return_lse=True)
        return out, lse, None
    return ops.patch_attention(qkv, cu_seqlens, max_seqlen, softmax_scale)
