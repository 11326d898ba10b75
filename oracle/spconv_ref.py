"""CPU restatement of spconv's sparse 3D convolution semantics (TEST ORACLE, PARITY UNPINNED).

spconv-cu124 (unpinned; reference environment.yml:47, scripts/build_image.sh:64) is not in
/root/reference and cannot be installed offline, and the reference holds no tests or golden
vectors for it.  This file restates its published semantics anchored on
the reference call sites:
  * SparseConvTensor(features, indices[N,4]=(b,x,y,z) int32, spatial_shape, batch_size)
        pointcept/models/utils/structure.py:139-146, sparse_unet/spconv_unet_v1m1_base.py:251-258
  * SubMConv3d  k=3 / k=5(padding=1 ignored) / k=1   ptv3m1:278-284,499-506; spunet:43-68,114-121,222-224
  * SparseConv3d k=2 s=2                              spunet:137-144
  * SparseInverseConv3d k=2 (same indice_key)         spunet:173-179
Conventions: kernel offset index k = (i0*K1 + i1)*K2 + i2; weight [Cout, K0, K1, K2, Cin];
output j reads input at coordinate  o*stride - padding + i*dilation;  SubM forces
padding = (K//2)*dilation, stride 1 and output set == input set in the same row order;
strided output set = unique reachable output coordinates in ascending linearised
(b,x,y,z) order (canonical form; any real spconv must be compared after sorting rows).
Rulebook canonical form: pair[KV, N_out] int32, input row or -1.
"""
import numpy as np
import torch


def _lin(idx, shape):
    idx = idx.astype(np.int64)
    return ((idx[:, 0] * shape[0] + idx[:, 1]) * shape[1] + idx[:, 2]) * shape[2] + idx[:, 3]


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(v)


def kernel_offsets(ksize):
    K0, K1, K2 = _triple(ksize)
    return [(i0, i1, i2) for i0 in range(K0) for i1 in range(K1) for i2 in range(K2)]


def _lookup(sorted_keys, sorted_rows, query):
    pos = np.searchsorted(sorted_keys, query)
    pos = np.minimum(pos, len(sorted_keys) - 1)
    hit = sorted_keys[pos] == query
    return np.where(hit, sorted_rows[pos], -1).astype(np.int32)


def subm_rulebook(indices, spatial_shape, ksize, dilation=1):
    """pair[KV, N]: pair[k, j] = row of the voxel at coord(j) + (i - K//2)*dilation, or -1."""
    indices = np.asarray(indices)
    shape = tuple(int(s) for s in spatial_shape)
    K = _triple(ksize)
    d = _triple(dilation)
    keys = _lin(indices, shape)
    srt = np.argsort(keys, kind="stable")
    sk, sr = keys[srt], srt.astype(np.int32)
    N = len(indices)
    pair = np.full((K[0] * K[1] * K[2], N), -1, dtype=np.int32)
    for k, off in enumerate(kernel_offsets(K)):
        nb = indices.astype(np.int64).copy()
        ok = np.ones(N, dtype=bool)
        for a in range(3):
            nb[:, 1 + a] += (off[a] - K[a] // 2) * d[a]
            ok &= (nb[:, 1 + a] >= 0) & (nb[:, 1 + a] < shape[a])
        res = _lookup(sk, sr, _lin(np.where(ok[:, None], nb, 0), shape))
        pair[k] = np.where(ok, res, -1)
    return pair


def conv_out_shape(spatial_shape, ksize, stride, padding, dilation):
    K, s, p, d = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    return [(int(spatial_shape[a]) + 2 * p[a] - d[a] * (K[a] - 1) - 1) // s[a] + 1 for a in range(3)]


def strided_rulebook(indices, spatial_shape, ksize, stride, padding=0, dilation=1):
    """-> out_indices[M,4] int32 (ascending key), out_shape, pair_fwd[KV,M] (input row per out row / offset),
    pair_bwd[KV,N] (out row per input row / offset)."""
    indices = np.asarray(indices)
    K, s, p, d = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    oshape = conv_out_shape(spatial_shape, ksize, stride, padding, dilation)
    N = len(indices)
    KV = K[0] * K[1] * K[2]
    cand_key = np.full((KV, N), -1, dtype=np.int64)
    cand = np.zeros((KV, N, 4), dtype=np.int64)
    for k, off in enumerate(kernel_offsets(K)):
        ok = np.ones(N, dtype=bool)
        o = indices.astype(np.int64).copy()
        for a in range(3):
            num = indices[:, 1 + a].astype(np.int64) + p[a] - off[a] * d[a]
            ok &= (num >= 0) & (num % s[a] == 0)
            oa = num // s[a]
            ok &= (oa >= 0) & (oa < oshape[a])
            o[:, 1 + a] = oa
        cand[k] = o
        cand_key[k] = np.where(ok, _lin(np.where(ok[:, None], o, 0), oshape), -1)
    valid = cand_key >= 0
    ukeys = np.unique(cand_key[valid])
    M = len(ukeys)
    pair_bwd = np.where(valid, np.searchsorted(ukeys, np.where(valid, cand_key, 0)), -1).astype(np.int32)
    pair_fwd = np.full((KV, M), -1, dtype=np.int32)
    for k in range(KV):
        rows = np.nonzero(valid[k])[0]
        pair_fwd[k, pair_bwd[k, rows]] = rows
    out_indices = np.zeros((M, 4), dtype=np.int32)
    rem = ukeys.copy()
    for a in (2, 1, 0):
        out_indices[:, 1 + a] = rem % oshape[a]
        rem //= oshape[a]
    out_indices[:, 0] = rem
    return out_indices, oshape, pair_fwd, pair_bwd


def conv_apply(features, weight, pair, bias=None):
    """out[j] = bias + sum_k features[pair[k,j]] @ weight[:,k,:].T  (torch, differentiable, fp32/fp64).

    weight [Cout, KV, Cin] (the module's [Cout,K0,K1,K2,Cin] flattened); pair [KV, N_out]."""
    KV, M = pair.shape
    out = features.new_zeros(M, weight.shape[0])
    pair_t = torch.as_tensor(pair, dtype=torch.long)
    for k in range(KV):
        rows = torch.nonzero(pair_t[k] >= 0).squeeze(1)
        if rows.numel() == 0:
            continue
        out = out.index_add(0, rows, features[pair_t[k, rows]] @ weight[:, k, :].t())
    if bias is not None:
        out = out + bias
    return out


def inverse_conv_apply(features, weight, pair_bwd, bias=None):
    """SparseInverseConv3d: out row i (original input set of the paired SparseConv3d) =
    sum_k features[pair_bwd[k,i]] @ weight[:,k,:].T  -- roles of in/out swapped, same offset index."""
    return conv_apply(features, weight, pair_bwd, bias)
