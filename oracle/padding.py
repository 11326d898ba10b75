"""numpy restatement of SerializedAttention.get_padding_and_inverse (TEST ORACLE).

Follows pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:114-170
and is pinned against tests/golden/padding.npz (generated from that code).
A scene with n > K points is padded to ceil(n/K)*K; the last patch is completed by
borrowing the K - n%K tokens that precede it (":144-154"); a scene with n <= K stays
one short sequence.
"""
import numpy as np


def padding_and_inverse(offset, K):
    offset = np.asarray(offset, dtype=np.int64)
    counts = np.diff(offset, prepend=0)
    padded = np.where(counts > K, (counts + K - 1) // K * K, counts)
    o = np.concatenate([[0], offset])
    op = np.concatenate([[0], np.cumsum(padded)])
    pad = np.arange(op[-1], dtype=np.int64)
    unpad = np.arange(o[-1], dtype=np.int64)
    cu = []
    for i in range(len(offset)):
        unpad[o[i]:o[i + 1]] += op[i] - o[i]
        if counts[i] != padded[i]:
            r = counts[i] % K
            pad[op[i + 1] - K + r: op[i + 1]] = pad[op[i + 1] - 2 * K + r: op[i + 1] - K]
        pad[op[i]:op[i + 1]] -= op[i] - o[i]
        cu.append(np.arange(op[i], op[i + 1], K, dtype=np.int32))
    cu_seqlens = np.concatenate(cu + [np.array([op[-1]], dtype=np.int32)]).astype(np.int32)
    return pad, unpad, cu_seqlens
