"""Independent corroboration of oracle/spconv_ref.py (the restatement of spconv's semantics): every sparse
convolution the hot path uses is compared against a DENSE torch convolution of the densified tensor --
F.conv3d for SubMConv3d / SparseConv3d, F.conv_transpose3d for SparseInverseConv3d -- forward and gradients.
This pins the offset <-> weight-slice mapping (k = (i0*K1 + i1)*K2 + i2, weight [Cout, K0, K1, K2, Cin],
cross-correlation, no flip), the strided convention (out o reads in o*s - p + i*d), the active-output rule
and the inverse-conv pairing against a definition that shares no code with the oracle.  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spconv_ref as osp

S = 16


def _voxels(p, seed, batch):
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(batch):
        occ = np.argwhere(rng.random((S, S, S)) < p)
        rows.append(np.concatenate([np.full((len(occ), 1), b), occ], 1))
    idx = np.concatenate(rows).astype(np.int32)
    return idx[rng.permutation(len(idx))]


def _densify(feat, idx, batch, shape):
    """[N, C] rows at (b, x, y, z) -> [B, C, X, Y, Z] (differentiable)."""
    dense = feat.new_zeros((batch, *shape, feat.shape[1]))
    i = torch.as_tensor(idx, dtype=torch.long)
    dense = dense.index_put((i[:, 0], i[:, 1], i[:, 2], i[:, 3]), feat)
    return dense.permute(0, 4, 1, 2, 3)


def _rows(dense, idx):
    i = torch.as_tensor(idx, dtype=torch.long)
    return dense.permute(0, 2, 3, 4, 1)[i[:, 0], i[:, 1], i[:, 2], i[:, 3]]


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(v)


def _check_grads(ref_out, out, dout, leaves_ref, leaves):
    gr = torch.autograd.grad(ref_out, leaves_ref, dout)
    go = torch.autograd.grad(out, leaves, dout)
    for a, b in zip(go, gr):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("ksize", [1, 3, 5, (3, 1, 5)])
@pytest.mark.parametrize("batch,p", [(1, 0.3), (2, 0.1)])
def test_subm_conv_equals_dense_conv3d_on_active_sites(ksize, batch, p):
    """SubMConv3d (ptv3m1:278-284,499-506; spunet:43-68,114-121,222-224): dense 'same' convolution read back at the
    active sites only; the module's padding argument is irrelevant."""
    torch.manual_seed(0)
    idx = _voxels(p, 1, batch)
    K = _triple(ksize)
    cin, cout, n = 5, 7, len(idx)
    feat = torch.randn(n, cin, dtype=torch.float64, requires_grad=True)
    w5 = torch.randn(cout, *K, cin, dtype=torch.float64, requires_grad=True)      # spconv parameter layout
    bias = torch.randn(cout, dtype=torch.float64, requires_grad=True)
    pair = osp.subm_rulebook(idx, [S, S, S], ksize)
    out = osp.conv_apply(feat, w5.reshape(cout, -1, cin), pair, bias)
    dense = F.conv3d(_densify(feat, idx, batch, (S, S, S)), w5.permute(0, 4, 1, 2, 3), bias, padding=tuple(k // 2 for k in K))
    ref = _rows(dense, idx)
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-11)
    _check_grads(ref, out, torch.randn_like(ref), (feat, w5, bias), (feat, w5, bias))
    # same rulebook in a larger (head-room) spatial shape, as Point.sparsify passes it (structure.py:137)
    assert np.array_equal(pair, osp.subm_rulebook(idx, [S + 96] * 3, ksize))


@pytest.mark.parametrize("ksize,stride,padding", [(2, 2, 0), (3, 2, 1), (3, 1, 1), (2, 1, 0)])
def test_strided_conv_and_inverse_equal_dense_conv3d_and_conv_transpose3d(ksize, stride, padding):
    """SparseConv3d k2 s2 (spunet:137-144) and its paired SparseInverseConv3d (spunet:173-179)."""
    torch.manual_seed(1)
    batch = 2
    idx = _voxels(0.15, 2, batch)
    cin, cout, n = 4, 6, len(idx)
    out_idx, oshape, pf, pb = osp.strided_rulebook(idx, [S, S, S], ksize, stride, padding)
    m = len(out_idx)
    feat = torch.randn(n, cin, dtype=torch.float64, requires_grad=True)
    w5 = torch.randn(cout, ksize, ksize, ksize, cin, dtype=torch.float64, requires_grad=True)
    out = osp.conv_apply(feat, w5.reshape(cout, -1, cin), pf)
    dense = F.conv3d(_densify(feat, idx, batch, (S, S, S)), w5.permute(0, 4, 1, 2, 3), None, stride=stride, padding=padding)
    assert list(dense.shape[2:]) == list(oshape)
    ref = _rows(dense, out_idx)
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-11)
    # active-output rule: the dense result is exactly zero outside the oracle's output set, and the set is ascending (b,x,y,z)
    mask = torch.ones(dense.shape[0], *dense.shape[2:], dtype=torch.bool)
    oi = torch.as_tensor(out_idx, dtype=torch.long)
    mask[oi[:, 0], oi[:, 1], oi[:, 2], oi[:, 3]] = False
    assert float(dense.detach().permute(0, 2, 3, 4, 1)[mask].abs().max() if mask.any() else 0.0) == 0.0
    key = ((oi[:, 0] * oshape[0] + oi[:, 1]) * oshape[1] + oi[:, 2]) * oshape[2] + oi[:, 3]
    assert bool((key[1:] > key[:-1]).all())
    _check_grads(ref, out, torch.randn_like(ref), (feat, w5), (feat, w5))
    # pair_fwd / pair_bwd describe the same pair set
    k, i = np.nonzero(pb >= 0)
    assert np.array_equal(pf[k, pb[k, i]], i) and (pf >= 0).sum() == (pb >= 0).sum()

    # inverse conv: rows of the M-set back to the N-set through the same rulebook = transposed dense convolution
    x = torch.randn(m, cout, dtype=torch.float64, requires_grad=True)
    wi5 = torch.randn(cin + 1, ksize, ksize, ksize, cout, dtype=torch.float64, requires_grad=True)   # [Cout', K.., Cin'=cout]
    outi = osp.inverse_conv_apply(x, wi5.reshape(cin + 1, -1, cout), pb)
    xd = _densify(x, out_idx, batch, tuple(oshape))
    natural = [(oshape[a] - 1) * stride - 2 * padding + ksize for a in range(3)]
    op = [S - natural[a] for a in range(3)]
    assert all(0 <= o < max(stride, 1) + 1 for o in op)
    densei = F.conv_transpose3d(xd, wi5.permute(4, 0, 1, 2, 3), None, stride=stride, padding=padding,
                                output_padding=tuple(min(o, stride - 1) for o in op) if stride > 1 else 0)
    if list(densei.shape[2:]) != [S, S, S]:       # stride 1 with a shrinking kernel: pad the missing border with zeros
        densei = F.pad(densei, (0, S - densei.shape[4], 0, S - densei.shape[3], 0, S - densei.shape[2]))
    refi = _rows(densei, idx)
    assert torch.allclose(outi, refi, rtol=1e-10, atol=1e-11)
    _check_grads(refi, outi, torch.randn_like(refi), (x, wi5), (x, wi5))


def test_subm_backward_flip_identity_holds_only_for_odd_kernels():
    """The product's backward-data pass reads the forward SubM table with flipped offsets: pair[k, j] = i <=>
    pair[KV-1-k, i] = j.  True for odd kernels (symmetric offset set); the drop-in refuses even SubM kernels."""
    idx = _voxels(0.3, 3, 2)
    for ks in (3, 5, (3, 1, 5)):
        pair = osp.subm_rulebook(idx, [S] * 3, ks)
        kv = pair.shape[0]
        k, j = np.nonzero(pair >= 0)
        assert np.array_equal(pair[kv - 1 - k, pair[k, j]], j)
