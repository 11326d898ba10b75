"""CPU restatement of the reference's GridSample transform and collate_fn -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows pointcept/datasets/transform.py:840-958 (GridSample.__call__), :980-995 (ravel_hash_vec), :997-1011 (fnv_hash_vec) and
pointcept/datasets/utils.py:19-73 (collate_fn: concatenate + cumulative offset).  Pinned against the reference's own class by
tests/golden/grid_sample.npz (tools/gen_golden.py gen_grid_sample): hashes, grid_coord, inverse, counts and the sampled
grid_coord are bit-exact.  One deliberate difference: the reference sorts with np.argsort's default (unstable) algorithm, so WHICH
member of a voxel lands at rank i is unspecified there; here the sort is stable and the fixtures only pin what the reference
determines (the voxel set, its order, inverse, counts, and that every picked point lies in its voxel).
"""
import numpy as np


def fnv_hash_vec(arr):
    arr = arr.astype(np.uint64, copy=True)
    h = np.full(arr.shape[0], np.uint64(14695981039346656037), dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(arr.shape[1]):
            h = h * np.uint64(1099511628211)
            h = np.bitwise_xor(h, arr[:, j])
    return h


def ravel_hash_vec(arr):
    arr = arr.copy()
    arr -= arr.min(0)
    arr = arr.astype(np.uint64, copy=False)
    arr_max = arr.max(0).astype(np.uint64) + np.uint64(1)
    keys = np.zeros(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1] - 1):
        keys += arr[:, j]
        keys *= arr_max[j + 1]
    keys += arr[:, -1]
    return keys


def cells(coord, grid_size, math="float64"):
    """floor(coord / grid_size) - min, min cell.  math = float64: NumPy >= 2 promotion of `float32 array / 0-d float64 array`;
    float32: NumPy 1.x value-based casting."""
    g = np.broadcast_to(np.asarray(grid_size, dtype=np.float64), (3,))
    if math == "float64":
        scaled = coord.astype(np.float64) / g
    else:
        scaled = coord.astype(np.float32) / g.astype(np.float32)
    grid = np.floor(scaled).astype(np.int64)
    mn = grid.min(0)
    scaled = scaled.copy()
    scaled -= mn          # in place, as transform.py:866: the result keeps scaled's dtype (float32 under NumPy 1.x semantics)
    return grid - mn, mn, scaled


def plan(coord, grid_size, hash_type="fnv", math="float64"):
    """-> dict(grid_coord [N,3], min_cell [3], key, idx_sort (stable), inverse [N], count [M], start [M])"""
    grid, mn, scaled = cells(coord, grid_size, math)
    key = fnv_hash_vec(grid) if hash_type == "fnv" else ravel_hash_vec(grid)
    idx_sort = np.argsort(key, kind="stable")
    key_sort = key[idx_sort]
    _, inv_sorted, count = np.unique(key_sort, return_inverse=True, return_counts=True)
    inverse = np.zeros_like(inv_sorted)
    inverse[idx_sort] = inv_sorted
    start = np.cumsum(np.insert(count, 0, 0)[0:-1])
    return dict(grid_coord=grid, min_cell=mn, scaled=scaled, key=key, idx_sort=idx_sort, inverse=inverse, count=count, start=start)


def select(p, member):
    """idx_unique for a per-voxel member rank array (train: u % count, test fragment i: i % count), transform.py:876-881,914-916"""
    return p["idx_sort"][p["start"] + member % p["count"]]


def batched(coords, grid_size, hash_type="fnv", math="float64"):
    """per-scene plans + the collated view (datasets/utils.py:19-73): concatenated inverse / grid_coord, cumulative voxel offsets"""
    plans = [plan(c, grid_size, hash_type, math) for c in coords]
    offset = np.cumsum([len(p["count"]) for p in plans])
    return plans, offset
