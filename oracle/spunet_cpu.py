"""CPU restatement of SpUNet-v1m1 (SpUNetBase.forward) on oracle/spconv_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows pointcept/models/sparse_unet/spconv_unet_v1m1_base.py: BasicBlock :23-85 (conv-bn-relu-conv-bn, projected residual when the
width changes, relu), stem :113-123 (SubMConv3d k5 + BN + ReLU, indice_key "stem"), stages :125-216 (SparseConv3d k2 s2 + BN + ReLU ->
blocks; SparseInverseConv3d k2 + BN + ReLU -> concat(up, skip) -> blocks), final SubMConv3d k1 :222-224, forward :244-280.
BatchNorm is BatchNorm1d(eps=1e-3, momentum=0.01) in training mode (:108).

Pinned by tests/test_oracle_model.py::test_spunet_restatement_matches_unmodified_reference_model: the UNMODIFIED reference class, run
on CPU over the same oracle convolutions (tools/ref_import.py), gives the same logits and parameter gradients.  The sparse-conv
arithmetic underneath stays "corroborated" (dense conv3d cross-check), as everywhere else.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import spconv_ref as osp


def forward(sd, b, model):
    """sd: state_dict (tensors; leave requires_grad on to differentiate), b: batch dict of numpy arrays (grid_coord, feat, offset),
    model: any object with ``num_stages``, ``layers`` and ``channels`` (the mirror or the reference class) -> logits [N, classes]"""
    bid = np.repeat(np.arange(len(b["offset"])), np.diff(b["offset"], prepend=0))
    idx = np.concatenate([bid[:, None], b["grid_coord"]], 1).astype(np.int32)
    shape = (b["grid_coord"].max(0) + 96).tolist()

    def bn(x, p):
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, 1e-3)

    def w(p):
        t = sd[p + ".weight"]
        return t.reshape(t.shape[0], -1, t.shape[-1])

    books = {}

    def subm(x, p, key, ks):
        if (key, ks) not in books:
            books[(key, ks)] = osp.subm_rulebook(levels[key][0], levels[key][1], ks)
        return osp.conv_apply(x, w(p), books[(key, ks)], sd.get(p + ".bias"))

    def block(x, p, key):
        res = x
        if p + ".proj.0.weight" in sd:
            res = bn(F.linear(x, w(p + ".proj.0")[:, 0, :]), p + ".proj.1")
        y = F.relu(bn(subm(x, p + ".conv1", key, 3), p + ".bn1"))
        y = bn(subm(y, p + ".conv2", key, 3), p + ".bn2")
        return F.relu(y + res)

    levels = {0: (idx, shape)}
    x = torch.from_numpy(b["feat"])
    x = F.relu(bn(subm(x, "conv_input.0", 0, 5), "conv_input.1"))
    skips, strided = [x], {}
    ns = model.num_stages
    for s in range(ns):
        oi, osh, pf, pb = osp.strided_rulebook(levels[s][0], levels[s][1], 2, 2)
        strided[s] = (pf, pb)
        levels[s + 1] = (oi, osh)
        x = F.relu(bn(osp.conv_apply(x, w(f"down.{s}.0"), pf), f"down.{s}.1"))
        for i in range(model.layers[s]):
            x = block(x, f"enc.{s}.block{i}", s + 1)
        skips.append(x)
    x = skips.pop(-1)
    for s in reversed(range(ns)):
        x = F.relu(bn(osp.inverse_conv_apply(x, w(f"up.{s}.0"), strided[s][1]), f"up.{s}.1"))
        x = torch.cat([x, skips.pop(-1)], 1)
        for i in range(model.layers[len(model.channels) - s - 1]):
            x = block(x, f"dec.{s}.block{i}", s)
    return F.linear(x, w("final")[:, 0, :], sd["final.bias"])
