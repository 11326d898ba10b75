"""Functional CPU (torch fp32) restatement of the PT-v3m1 forward pass (TEST ORACLE + CPU baseline).

Follows pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:
  Embedding :485-515, Block :318-338, SerializedAttention :172-222 (dense math of :190-206 per patch,
  padding :114-170), SerializedPooling :371-444, SerializedUnpooling :471-482, PointTransformerV3.forward
  :699-714, and Point.serialization / sparsify (models/utils/structure.py:53-148), with the sparse
  convolutions restated by oracle/spconv_ref.py.  Weights come from a state_dict with the reference's
  parameter names.  shuffle_orders must be False and DropPath inactive (eval / drop_path = 0) for parity runs.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import attention as oattn
from . import padding as opad
from . import serialization as oser
from . import spconv_ref as osp


# tiny PT-v3m1 used by tests/golden/ptv3_tiny.npz (tools/gen_golden.py) -- head dim 16 as in every stock config
TINY_CFG = dict(
    in_channels=6, order=("z", "z-trans", "hilbert", "hilbert-trans"), stride=(2, 2), enc_depths=(2, 2, 2),
    enc_channels=(16, 32, 64), enc_num_head=(1, 2, 4), enc_patch_size=(128, 32, 8), dec_depths=(2, 2),
    dec_channels=(32, 32), dec_num_head=(2, 2), dec_patch_size=(128, 32), mlp_ratio=4, drop_path=0.0,
    shuffle_orders=False,
)


def _bn(x, sd, prefix, training, eps=1e-3):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        return F.batch_norm(x, None, None, w, b, True, 0.0, eps)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b, False, 0.0, eps)


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


class _Pt:
    """plain attribute bag for one resolution level"""


def _serialize(pt, orders):
    code, order, inverse, depth = oser.serialize(pt.grid, pt.batch, orders, pt.depth)
    pt.code, pt.order, pt.inverse = code, order, inverse


def _rulebook(pt, ksize):
    key = ("subm", ksize)
    if key not in pt.rulebooks:
        idx = np.concatenate([pt.batch[:, None], pt.grid], 1).astype(np.int32)
        pt.rulebooks[key] = osp.subm_rulebook(idx, pt.sparse_shape, ksize)
    return pt.rulebooks[key]


def _subm(x, pt, sd, prefix, ksize):
    w = sd[prefix + ".weight"]
    w = w.reshape(w.shape[0], -1, w.shape[-1])
    return osp.conv_apply(x, w, _rulebook(pt, ksize), sd.get(prefix + ".bias"))


def _attention(x, pt, sd, prefix, heads, K, order_index, attn_dtype=None):
    C = x.shape[1]
    if "pad" not in pt.cache:
        pt.cache["pad"] = opad.padding_and_inverse(pt.offset, K)
    pad, unpad, cu = pt.cache["pad"]
    order = torch.from_numpy(pt.order[order_index][pad])
    inverse = torch.from_numpy(unpad[pt.inverse[order_index]])
    qkv = _lin(x, sd, prefix + ".qkv")[order]
    if attn_dtype is not None:  # emulate the bf16 cast the reference applies before flash-attn (:209)
        qkv = qkv.to(attn_dtype).float()
    out = oattn.varlen_attention(qkv.reshape(-1, 3, heads, C // heads), cu, (C // heads) ** -0.5).reshape(-1, C)
    if attn_dtype is not None:
        out = out.to(attn_dtype).float()
    return _lin(out[inverse], sd, prefix + ".proj")


def _block(x, pt, sd, prefix, heads, K, order_index, attn_dtype):
    y = _subm(x, pt, sd, prefix + ".cpe.0", 3)
    y = _ln(_lin(y, sd, prefix + ".cpe.1"), sd, prefix + ".cpe.2")
    x = x + y
    x = x + _attention(_ln(x, sd, prefix + ".norm1.0"), pt, sd, prefix + ".attn", heads, K, order_index, attn_dtype)
    h = _ln(x, sd, prefix + ".norm2.0")
    h = _lin(F.gelu(_lin(h, sd, prefix + ".mlp.0.fc1")), sd, prefix + ".mlp.0.fc2")
    return x + h


def _pool(x, pt, sd, prefix, orders, training):
    code = pt.code >> 3
    uniq, cluster = np.unique(code[0], return_inverse=True)
    m = len(uniq)
    idx = np.argsort(cluster, kind="stable")
    counts = np.bincount(cluster, minlength=m)
    ptr = np.concatenate([[0], np.cumsum(counts)])
    head = idx[ptr[:-1]]
    proj = _lin(x, sd, prefix + ".proj")
    cl = torch.from_numpy(cluster)
    feat = torch.full((m, proj.shape[1]), -float("inf")).scatter_reduce(0, cl[:, None].expand(-1, proj.shape[1]), proj, "amax",
                                                                          include_self=True)
    nxt = _Pt()
    nxt.grid = pt.grid[head] >> 1
    nxt.batch = pt.batch[head]
    nxt.depth = pt.depth - 1
    nxt.offset = np.cumsum(np.bincount(nxt.batch, minlength=len(pt.offset)))
    nxt.code = code[:, head]
    nxt.order = np.argsort(nxt.code, axis=1, kind="stable")
    nxt.inverse = np.empty_like(nxt.order)
    for r in range(nxt.order.shape[0]):
        nxt.inverse[r, nxt.order[r]] = np.arange(m)
    nxt.sparse_shape = [int(v) + 96 for v in nxt.grid.max(0)]
    nxt.rulebooks, nxt.cache = {}, {}
    nxt.cluster, nxt.parent = cl, pt
    feat = F.gelu(_bn(feat, sd, prefix + ".norm.0", training))
    return feat, nxt


def forward(sd, data, cfg, bn_training=True, attn_dtype=None):
    """sd: state_dict-like mapping of (fp32, CPU) tensors with the PT-v3m1 names; data: grid_coord [N,3] int,
    feat [N,Cin] float32, offset [B] (numpy or torch).  Returns final per-point features [N, dec_channels[0]]."""
    sd = {k: v for k, v in sd.items()}
    orders = list(cfg["order"])
    pt = _Pt()
    pt.grid = np.asarray(data["grid_coord"]).astype(np.int64)
    pt.offset = np.asarray(data["offset"]).astype(np.int64)
    pt.batch = np.repeat(np.arange(len(pt.offset)), np.diff(pt.offset, prepend=0))
    pt.depth = oser.serialization_depth(pt.grid)
    _serialize(pt, orders)
    pt.sparse_shape = [int(v) + 96 for v in pt.grid.max(0)]
    pt.rulebooks, pt.cache = {}, {}
    x = torch.as_tensor(data["feat"]).float()
    x = _subm(x, pt, sd, "embedding.stem.conv", 5)
    x = F.gelu(_bn(x, sd, "embedding.stem.norm", bn_training))
    n_stage = len(cfg["enc_depths"])
    feats = {}
    for s in range(n_stage):
        if s > 0:
            x, pt = _pool(x, pt, sd, f"enc.enc{s}.down", orders, bn_training)
        for i in range(cfg["enc_depths"][s]):
            x = _block(x, pt, sd, f"enc.enc{s}.block{i}", cfg["enc_num_head"][s], cfg["enc_patch_size"][s], i % len(orders), attn_dtype)
        feats[s] = (x, pt)
    for s in reversed(range(n_stage - 1)):
        skip, parent = feats[s]
        up = F.gelu(_bn(_lin(x, sd, f"dec.dec{s}.up.proj.0"), sd, f"dec.dec{s}.up.proj.1", bn_training))
        sk = F.gelu(_bn(_lin(skip, sd, f"dec.dec{s}.up.proj_skip.0"), sd, f"dec.dec{s}.up.proj_skip.1", bn_training))
        # m1 quirk (ptv3m1:471-482): the first decoder block's CPE conv still sees proj_skip(parent) only
        x_conv_in = sk
        x = sk + up[pt.cluster]
        pt = parent
        for i in range(cfg["dec_depths"][s]):
            prefix = f"dec.dec{s}.block{i}"
            if i == 0:
                y = _subm(x_conv_in, pt, sd, prefix + ".cpe.0", 3)
                y = _ln(_lin(y, sd, prefix + ".cpe.1"), sd, prefix + ".cpe.2")
                x = x + y
                x = x + _attention(_ln(x, sd, prefix + ".norm1.0"), pt, sd, prefix + ".attn", cfg["dec_num_head"][s],
                                   cfg["dec_patch_size"][s], 0, attn_dtype)
                h = _ln(x, sd, prefix + ".norm2.0")
                x = x + _lin(F.gelu(_lin(h, sd, prefix + ".mlp.0.fc1")), sd, prefix + ".mlp.0.fc2")
            else:
                x = _block(x, pt, sd, prefix, cfg["dec_num_head"][s], cfg["dec_patch_size"][s], i % len(orders), attn_dtype)
    return x
