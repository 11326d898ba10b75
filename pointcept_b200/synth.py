"""Seeded synthetic scenes (host, numpy) shaped like the reference's collated batches
(pointcept/datasets/utils.py:19-73: flat tensors + cumulative ``offset``).

indoor: ScanNet-like room at 2 cm voxels -- floor, ceiling-less walls and axis-aligned boxes sampled on
their surfaces, voxelised and de-duplicated, calibrated to ~11 active 3^3 neighbours per voxel
(SURVEY.md section 8(d)).  lidar: nuScenes-like 5 cm sweep (polar rings on a ground plane + a few objects).
"""
import numpy as np


def _surface_points(rng, lo, hi, n):
    """n points uniformly on the 6 faces of the box [lo, hi]."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    ext = hi - lo
    areas = np.array([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2], ext[0] * ext[1], ext[0] * ext[1]])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    p = lo + rng.random((n, 3)) * ext
    ax = face // 2
    p[np.arange(n), ax] = np.where(face % 2 == 0, lo[ax], hi[ax])
    return p


def indoor_scene(seed, target_voxels=120_000, grid=0.02, jitter=0.15):
    rng = np.random.default_rng(seed)
    # room ~ 8 x 6 x 2.6 m scaled so that the surface area yields ~target voxels at `grid`
    scale = np.sqrt(target_voxels / 120_000.0)
    room = np.array([8.0, 6.0, 2.6]) * np.array([scale, scale, 1.0])
    pts = []
    dens = 1.8 / grid ** 2  # samples per m^2, calibrated (with jitter) to ~11 active 3^3 neighbours per voxel
    floor_n = int(room[0] * room[1] * dens)
    f = rng.random((floor_n, 3)) * room
    f[:, 2] = 0.0
    pts.append(f)
    for ax, val in ((0, 0.0), (0, room[0]), (1, 0.0), (1, room[1])):
        other = 1 - ax
        w = rng.random((int(room[other] * room[2] * dens), 3)) * room
        w[:, ax] = val
        pts.append(w)
    n_box = int(14 * scale * scale)
    for _ in range(n_box):
        size = rng.uniform([0.4, 0.4, 0.3], [1.8, 1.2, 1.1])
        lo = np.concatenate([rng.uniform([0.2, 0.2], room[:2] - size[:2] - 0.2), [0.0]])
        area = 2 * (size[0] * size[1] + size[1] * size[2] + size[0] * size[2])
        pts.append(_surface_points(rng, lo, lo + size, int(area * dens)))
    p = np.concatenate(pts)
    p = p + rng.normal(0.0, jitter * grid, p.shape)  # sensor noise: gives surfaces some thickness
    g = np.floor(p / grid).astype(np.int64)
    g -= g.min(0)
    _, first = np.unique(g, axis=0, return_index=True)
    first = np.sort(first)
    rng.shuffle(first)
    if len(first) > target_voxels:
        # crop like SphereCrop(point_max): keep the target_voxels nearest to a random centre
        c = p[first[rng.integers(len(first))]]
        d = ((p[first] - c) ** 2).sum(1)
        first = first[np.argsort(d)[:target_voxels]]
    g = g[first]
    g -= g.min(0)
    coord = (g + 0.5) * grid
    return coord.astype(np.float32), g.astype(np.int32)


def lidar_scene(seed, target_voxels=300_000, grid=0.05):
    rng = np.random.default_rng(seed)
    n_ring, n_az = 64, int(target_voxels / 64 * 1.6)
    elev = np.deg2rad(np.linspace(-25, 3, n_ring))
    az = rng.random(n_az) * 2 * np.pi
    E, A = np.meshgrid(elev, az, indexing="ij")
    h = 1.8
    r = np.where(E < -0.01, h / np.tan(-E), 60.0)
    r = np.minimum(r, 60.0) * (1 + 0.01 * rng.standard_normal(r.shape))
    x, y = r * np.cos(A), r * np.sin(A)
    z = np.where(E < -0.01, -h + 0.02 * rng.standard_normal(r.shape), r * np.tan(E))
    p = np.stack([x.ravel(), y.ravel(), z.ravel()], 1)
    g = np.floor(p / grid).astype(np.int64)
    g -= g.min(0)
    _, first = np.unique(g, axis=0, return_index=True)
    if len(first) > target_voxels:
        first = rng.choice(first, target_voxels, replace=False)
    g = g[first]
    g -= g.min(0)
    return ((g + 0.5) * grid).astype(np.float32), g.astype(np.int32)


def make_batch(n_scenes, seed=0, kind="indoor", target_voxels=None, in_channels=None, num_classes=20):
    """-> dict of numpy arrays shaped like a collated Pointcept batch."""
    coords, grids, feats, segs, offset = [], [], [], [], []
    tot = 0
    for s in range(n_scenes):
        if kind == "indoor":
            c, g = indoor_scene(seed * 1000 + s, target_voxels or 120_000)
            ch = in_channels or 6
        else:
            c, g = lidar_scene(seed * 1000 + s, target_voxels or 300_000)
            ch = in_channels or 4
        rng = np.random.default_rng(seed * 1000 + s + 7)
        coords.append(c)
        grids.append(g)
        feats.append(rng.standard_normal((len(c), ch)).astype(np.float32))
        segs.append(rng.integers(0, num_classes, len(c)).astype(np.int64))
        tot += len(c)
        offset.append(tot)
    return dict(coord=np.concatenate(coords), grid_coord=np.concatenate(grids), feat=np.concatenate(feats),
                segment=np.concatenate(segs), offset=np.array(offset, dtype=np.int64))


def neighbour_stats(grid_coord):
    """mean number of active 3^3 neighbours per voxel (including the centre)."""
    g = grid_coord.astype(np.int64)
    S = g.max(0) + 3
    key = ((g[:, 0] + 1) * S[1] + g[:, 1] + 1) * S[2] + g[:, 2] + 1
    ks = np.sort(key)
    tot = 0
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                q = key + (dx * S[1] + dy) * S[2] + dz
                pos = np.minimum(np.searchsorted(ks, q), len(ks) - 1)
                tot += (ks[pos] == q).sum()
    return tot / len(g)
