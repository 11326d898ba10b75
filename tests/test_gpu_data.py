"""GPU voxelisation + collate through the C ABI against the oracle restatement of the reference's GridSample /
collate_fn and against the fixture its own class produced.  Integer results are bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import grid_sample as ogs
from pointcept_b200 import datasets, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scenes(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_sample.npz"))
    return g, [g["coord0"], g["coord1"]]


def _raw_batch(scenes, rng):
    per = []
    for c in scenes:
        n = len(c)
        per.append(dict(coord=torch.from_numpy(c).to(DEV), color=torch.from_numpy(rng.random((n, 3)).astype(np.float32)).to(DEV),
                        normal=torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)).to(DEV),
                        segment=torch.from_numpy(rng.integers(-1, 20, n).astype(np.int32)).to(DEV)))
    return per


@pytest.mark.parametrize("hash_type", ["fnv", "ravel"])
@pytest.mark.parametrize("gs", [0.05, 0.02])
def test_batched_plan_bit_exact_vs_reference_fixture(golden_dir, hash_type, gs):
    g, scenes = _scenes(golden_dir)
    coord = torch.from_numpy(np.concatenate(scenes)).to(DEV)
    off = np.cumsum([len(c) for c in scenes]).tolist()
    plan = ops.grid_sample_plan(coord, off, gs, hash_type, "float64")
    idx0 = plan.select("test", 0).cpu().numpy()
    last = plan.select("test", max(plan.count_max_host) - 1).cpu().numpy()
    inverse, grid = plan.inverse.cpu().numpy(), plan.grid_coord.cpu().numpy()
    start, vstart = 0, 0
    for i, c in enumerate(scenes):
        tag = f"{hash_type}_{gs}_{i}"
        end, vend = off[i], plan.new_offset_host[i]
        assert np.array_equal(inverse[start:end], g[tag + "_inverse"]), tag
        assert vend - vstart == len(g[tag + "_grid_coord"])
        assert np.array_equal(grid[idx0[vstart:vend]], g[tag + "_grid_coord"]), tag          # voxel set AND output order
        assert plan.count_max_host[i] == int(g[tag + "_n_fragments"])
        assert np.array_equal(np.asarray(plan.min_cell_host[i]) * np.float64(gs), g[tag + "_min_coord"][0])
        # the stable member order: exactly the oracle's picks, for the first and the last fragment
        p = ogs.plan(c, gs, hash_type, "float64")
        assert np.array_equal(idx0[vstart:vend] - start, ogs.select(p, 0))
        assert np.array_equal(last[vstart:vend] - start, ogs.select(p, max(plan.count_max_host) - 1))
        assert np.array_equal(plan.vox_count.cpu().numpy()[vstart:vend], p["count"])
        assert np.array_equal(plan.sort_index.cpu().numpy()[start:end] - start, p["idx_sort"])
        start, vstart = end, vend


@pytest.mark.parametrize("math", ["float64", "float32"])
def test_transform_test_mode_and_collate(golden_dir, math):
    g, scenes = _scenes(golden_dir)
    rng = np.random.default_rng(0)
    per = _raw_batch(scenes, rng)
    batch = datasets.collate_fn(per)
    assert batch["offset"].tolist() == np.cumsum([len(c) for c in scenes]).tolist()
    tr = datasets.GridSample(grid_size=0.05, mode="test", return_inverse=True, return_grid_coord=True, return_min_coord=True,
                             return_displacement=True, math=math)
    parts = tr(dict(batch))
    plans = [ogs.plan(c, 0.05, "fnv", math) for c in scenes]
    assert len(parts) == max(int(p["count"].max()) for p in plans)
    covered = np.zeros(sum(len(c) for c in scenes), dtype=bool)
    for f, part in enumerate(parts):
        want_idx = np.concatenate([ogs.select(p, f) + s for p, s in zip(plans, [0, len(scenes[0])])])
        assert np.array_equal(part["index"].cpu().numpy(), want_idx)
        covered[want_idx] = True
        for key in ("coord", "color", "normal", "segment"):
            assert torch.equal(part[key], batch[key][part["index"]]), key
        assert np.array_equal(part["grid_coord"].cpu().numpy(), np.concatenate([p["grid_coord"] for p in plans])[want_idx])
        assert part["offset"].tolist() == np.cumsum([len(p["count"]) for p in plans]).tolist()
        disp = np.concatenate([p["scaled"] - p["grid_coord"] - 0.5 for p in plans])[want_idx]
        got = part["displacement"].cpu().numpy()
        assert got.dtype == disp.dtype and np.array_equal(got, disp)
        assert np.array_equal(part["min_coord"].cpu().numpy(), np.stack([p["min_cell"] * np.float64(0.05) for p in plans]))
    assert covered.all()     # the fragments of test mode cover every raw point (transform.py:913-916)
    assert np.array_equal(parts[0]["inverse"].cpu().numpy(), np.concatenate([p["inverse"] for p in plans]))


def test_transform_train_mode_properties(golden_dir):
    g, scenes = _scenes(golden_dir)
    rng = np.random.default_rng(1)
    batch = datasets.collate_fn(_raw_batch(scenes, rng))
    plans = [ogs.plan(c, 0.05, "fnv", "float64") for c in scenes]
    n_vox = sum(len(p["count"]) for p in plans)
    inv_all = np.concatenate([p["inverse"] + o for p, o in zip(plans, [0, len(plans[0]["count"])])])
    picks = []
    for seed in (3, 4):
        out = datasets.GridSample(grid_size=0.05, mode="train", return_grid_coord=True, seed=seed)(dict(batch))
        assert out["coord"].shape[0] == n_vox and out["offset"].tolist() == np.cumsum([len(p["count"]) for p in plans]).tolist()
        # exactly one point per voxel, voxels in the reference's order (ascending hash inside each scene)
        got = out["coord"].cpu().numpy()
        grid_want = np.concatenate([p["grid_coord"][ogs.select(p, 0)] for p in plans])
        assert np.array_equal(out["grid_coord"].cpu().numpy(), grid_want)
        picks.append(got)
    assert not np.array_equal(picks[0], picks[1])      # different seeds pick different members somewhere
    again = datasets.GridSample(grid_size=0.05, mode="train", return_grid_coord=True, seed=3)(dict(batch))
    assert np.array_equal(again["coord"].cpu().numpy(), picks[0])   # counter-based generator: same seed, same picks
    # member choice statistics: over many seeds every member of a 2-point voxel is chosen
    plan = ops.grid_sample_plan(batch["coord"], batch["offset"].tolist(), 0.05)
    cnt = plan.vox_count.cpu().numpy()
    two = np.where(cnt == 2)[0][:50]
    seen = [set() for _ in two]
    for seed in range(24):
        idx = plan.select("train", seed).cpu().numpy()
        assert np.array_equal(inv_all[idx], np.arange(n_vox))       # every pick lies in its own voxel
        for s, v in zip(seen, two):
            s.add(int(idx[v]))
    assert all(len(s) == 2 for s in seen)


def test_gather_rows_any_payload():
    rng = np.random.default_rng(2)
    idx = torch.from_numpy(rng.integers(0, 1000, 777)).to(DEV)
    for shape, dt in (((1000, 3), torch.float32), ((1000,), torch.int32), ((1000, 6), torch.float16), ((1000, 4), torch.int64),
                      ((1000, 3), torch.uint8), ((1000, 5), torch.float64), ((1000, 2, 3), torch.int16)):
        src = torch.from_numpy(rng.integers(0, 100, shape)).to(DEV).to(dt)
        assert torch.equal(ops.gather_rows(src, idx), src[idx]), (shape, dt)
    assert ops.gather_rows(src, idx[:0]).shape == (0, 2, 3)


def test_scene_layout_round_trip(tmp_path, golden_dir):
    g, scenes = _scenes(golden_dir)
    rng = np.random.default_rng(3)
    dirs = []
    for i, c in enumerate(scenes):
        d = tmp_path / "train" / f"scene{i:04d}_00"
        datasets.write_scene(str(d), c, color=rng.integers(0, 256, c.shape), normal=rng.normal(size=c.shape), segment=rng.integers(-1, 20, len(c)))
        dirs.append(str(d))
    loaded = [datasets.load_scene(d) for d in dirs]
    assert loaded[0]["name"] == "scene0000_00" and loaded[0]["split"] == "train"
    assert loaded[0]["color"].dtype == torch.float32 and loaded[0]["segment"].dtype == torch.int32
    assert (loaded[1]["instance"] == -1).all()
    batch = datasets.collate_fn([{k: v for k, v in s.items() if isinstance(v, torch.Tensor)} for s in loaded])
    out = datasets.GridSample(grid_size=0.02, mode="train", return_grid_coord=True, seed=0)(batch)
    want = sum(len(g[f"fnv_0.02_{i}_grid_coord"]) for i in range(2))
    assert out["grid_coord"].shape == (want, 3) and out["color"].shape == (want, 3) and out["segment"].shape == (want,)


def test_large_batch_matches_oracle():
    """8 scenes of ~200 k raw points at ScanNet's 2 cm grid: bit-exact inverse / order / counts against the numpy oracle."""
    rng = np.random.default_rng(5)
    scenes = []
    for i in range(8):
        n = 150000 + 20000 * i
        c = rng.random((n, 3)) * np.asarray([6.0, 5.0, 2.5]) - 3.0
        w = rng.integers(0, 3, n)
        c[np.arange(n), w] = np.round(c[np.arange(n), w]) + rng.normal(0, 0.003, n)
        scenes.append(c.astype(np.float32))
    off = np.cumsum([len(c) for c in scenes]).tolist()
    plan = ops.grid_sample_plan(torch.from_numpy(np.concatenate(scenes)).to(DEV), off, 0.02)
    inv = plan.inverse.cpu().numpy()
    idx0 = plan.select("test", 0).cpu().numpy()
    s = v = 0
    for i, c in enumerate(scenes):
        p = ogs.plan(c, 0.02)
        assert np.array_equal(inv[s:off[i]], p["inverse"])
        assert np.array_equal(idx0[v:plan.new_offset_host[i]] - s, ogs.select(p, 0))
        s, v = off[i], plan.new_offset_host[i]
