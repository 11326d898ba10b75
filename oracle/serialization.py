"""numpy restatement of the reference's point serialization (TEST ORACLE).

Follows, and is pinned bit-exactly against (tests/golden/serialization.npz):
  * pointcept/models/utils/serialization/z_order.py:39-49,66-101  (Morton key,
    x is the most significant bit of every triple)
  * pointcept/models/utils/serialization/hilbert.py:91-192        (Skilling
    transform, MSB-first over bits then dims, Gray->binary over the interleaved
    3*depth bit string)
  * pointcept/models/utils/serialization/default.py:9-24          (order
    dispatch, "-trans" swaps x and y, batch goes above bit 3*depth)
  * pointcept/models/utils/structure.py:85-100                    (argsort of
    the stacked codes and the inverse permutation)
"""
import numpy as np

ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")


def z_order_key(x, y, z, depth):
    x = x.astype(np.int64); y = y.astype(np.int64); z = z.astype(np.int64)
    key = np.zeros_like(x)
    for i in range(depth):
        bit = np.int64(1) << i
        key |= ((x & bit) << (2 * i + 2)) | ((y & bit) << (2 * i + 1)) | ((z & bit) << (2 * i))
    return key


def hilbert_key(x, y, z, depth):
    """Skilling's axes->transpose on three integer lanes, then interleave + prefix-xor."""
    X = [x.astype(np.int64).copy(), y.astype(np.int64).copy(), z.astype(np.int64).copy()]
    lowmask = (np.int64(1) << depth) - 1
    X = [v & lowmask for v in X]
    # hilbert.py:150-169 -- bit = 0 is the MSB (weight 2^(depth-1)); "lower bits" = weights below it
    for b in range(depth):
        q = np.int64(1) << (depth - 1 - b)
        p = q - 1
        for d in range(3):
            on = (X[d] & q) != 0
            # bit set: invert the lower bits of lane 0
            X[0] = np.where(on, X[0] ^ p, X[0])
            # bit clear: exchange lower bits of lane 0 and lane d where they differ
            t = np.where(on, 0, (X[0] ^ X[d]) & p)
            X[d] = X[d] ^ t
            X[0] = X[0] ^ t
    # interleave: per bit (MSB first) lanes 0,1,2 -> lane 0 is the top bit of each triple
    g = np.zeros_like(X[0])
    for i in range(depth):
        bit = np.int64(1) << i
        g |= ((X[0] & bit) << (2 * i + 2)) | ((X[1] & bit) << (2 * i + 1)) | ((X[2] & bit) << (2 * i))
    # Gray -> binary over the whole string (hilbert.py:69-88,175)
    s = 1
    while s < 3 * depth:
        g ^= g >> s
        s *= 2
    return g


def encode(grid_coord, batch, depth, order):
    """default.py:9-24.  grid_coord [N,3] int, batch [N] int or None -> int64 [N]."""
    assert order in ORDERS
    gc = np.asarray(grid_coord)
    x, y, z = gc[:, 0], gc[:, 1], gc[:, 2]
    if order.endswith("-trans"):
        x, y = y, x
    code = z_order_key(x, y, z, depth) if order.startswith("z") else hilbert_key(x, y, z, depth)
    if batch is not None:
        code = (np.asarray(batch).astype(np.int64) << (3 * depth)) | code
    return code


def serialization_depth(grid_coord):
    """structure.py:74."""
    return int(int(np.asarray(grid_coord).max()) + 1).bit_length()


def serialize(grid_coord, batch, orders, depth=None):
    """structure.py:72-100 (without the random order shuffle).  Returns code, order, inverse [k,N]."""
    if depth is None:
        depth = serialization_depth(grid_coord)
    code = np.stack([encode(grid_coord, batch, depth, o) for o in orders])
    order = np.argsort(code, axis=1, kind="stable")
    inverse = np.empty_like(order)
    n = code.shape[1]
    for r in range(code.shape[0]):
        inverse[r, order[r]] = np.arange(n)
    return code, order, inverse, depth
