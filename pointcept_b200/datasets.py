"""GPU voxelisation + collate: the step in front of the hot path.

Mirrors, for CUDA tensors and whole batches, the reference's

* ``GridSample`` transform      pointcept/datasets/transform.py:840-958  (same constructor arguments, same output keys)
* ``collate_fn``                pointcept/datasets/utils.py:19-73        (concatenate per-point keys, cumulative ``offset``)
* on-disk scene layout          pointcept/datasets/defaults.py:35-43,102-143 / scannet.py:29-35: one directory per scene with
                                ``coord.npy  color.npy  normal.npy  segment20.npy  instance.npy``

The reference voxelises every scene on a CPU worker with numpy (hash, argsort, unique) and concatenates afterwards; here the raw
scenes are concatenated first (``collate_fn``) and ``GridSample`` voxelises the whole batch with one set of launches
(csrc/voxelize.cuh) -- the result is the collated, voxelised batch.  No CPU fallback: tensors must be CUDA tensors.
"""
import os
from collections.abc import Mapping, Sequence

import numpy as np
import torch

from . import ops

DEFAULT_INDEX_VALID_KEYS = ["coord", "color", "normal", "superpoint", "strength", "segment", "instance"]   # transform.py:27-35
VALID_ASSETS = ["coord", "color", "normal", "strength", "segment", "instance", "pose"]                    # defaults.py:35-43


def index_operator(data_dict, index, duplicate=False):
    """transform.py:23-50: select rows of every key listed in ``index_valid_keys``."""
    if "index_valid_keys" not in data_dict:
        data_dict["index_valid_keys"] = list(DEFAULT_INDEX_VALID_KEYS)
    if not duplicate:
        for key in data_dict["index_valid_keys"]:
            if key in data_dict:
                data_dict[key] = ops.gather_rows(data_dict[key], index)
        return data_dict
    out = dict(index_valid_keys=data_dict["index_valid_keys"])
    for key in data_dict.keys():
        if key in data_dict["index_valid_keys"]:
            out[key] = ops.gather_rows(data_dict[key], index)
        elif key != "index_valid_keys":
            out[key] = data_dict[key]
    return out


def collate_fn(batch):
    """datasets/utils.py:19-73 for lists of tensors / dicts of tensors: per-point keys are concatenated, keys containing
    "offset" become cumulative, a dict without one gets ``offset`` from its ``coord`` lengths (what the reference's datasets
    add through the ``Collect`` transform, transform.py:75-101)."""
    if not isinstance(batch, Sequence):
        raise TypeError(f"{type(batch)} is not supported.")
    first = batch[0]
    if isinstance(first, torch.Tensor):
        return torch.cat(list(batch))
    if isinstance(first, str):
        return list(batch)
    if isinstance(first, list):                      # python lists of numbers (utils.py:32-34)
        return torch.cat([torch.tensor(d) for d in batch])
    if isinstance(first, Mapping):
        out = {}
        for key in first:
            if key == "index_valid_keys":
                out[key] = list(first[key])
            elif "offset" in key:
                out[key] = torch.cumsum(torch.cat([torch.diff(d[key], prepend=d[key].new_zeros(1)) for d in batch]), dim=0)
            else:
                out[key] = collate_fn([d[key] for d in batch])
        if "offset" not in out and "coord" in first:
            sizes = torch.tensor([d["coord"].shape[0] for d in batch], dtype=torch.int64)
            out["offset"] = torch.cumsum(sizes, 0).to(first["coord"].device)
        return out
    if isinstance(first, Sequence):                  # tuples of per-point tensors: the reference appends the cumulative offset
        cols = [collate_fn(list(s)) for s in zip(*batch)]         # (int32, from the first entry's lengths) as the last item (utils.py:35-40)
        sizes = torch.tensor([d[0].shape[0] for d in batch], dtype=torch.int64, device=cols[0].device)
        cols.append(torch.cumsum(sizes, dim=0).int())
        return cols
    return torch.utils.data.dataloader.default_collate(batch)


class GridSample:
    """Batched CUDA GridSample with the reference's constructor (transform.py:841-862).

    ``data_dict``: ``coord [N,3]`` fp32 plus any per-point keys, optionally ``offset`` (cumulative scene sizes; absent = one
    scene).  Integer outputs (``grid_coord``, ``inverse``, the voxel order, counts, ``min_coord``) equal the reference's;
    the member picked inside a voxel follows the reference's rule with a counter-based generator in train mode (seeded through
    ``seed`` or torch's global generator) and the stable sort order in test mode (np.argsort's default order among equal keys
    is unspecified in the reference).  ``math``: "float64" = NumPy >= 2 promotion of ``coord / np.array(grid_size)``,
    "float32" = NumPy 1.x.
    """

    def __init__(self, grid_size=0.05, hash_type="fnv", mode="train", return_inverse=False, return_grid_coord=False,
                 return_min_coord=False, return_displacement=False, project_displacement=False, math="float64", seed=None):
        assert mode in ["train", "test"]
        assert hash_type in ["fnv", "ravel"]
        self.grid_size, self.hash_type, self.mode = grid_size, hash_type, mode
        self.return_inverse, self.return_grid_coord, self.return_min_coord = return_inverse, return_grid_coord, return_min_coord
        self.return_displacement, self.project_displacement = return_displacement, project_displacement
        self.math, self.seed = math, seed

    def _offsets(self, data_dict):
        n = data_dict["coord"].shape[0]
        if "offset" in data_dict:
            off = data_dict["offset"]
            return [int(x) for x in (off.tolist() if isinstance(off, torch.Tensor) else off)]
        return [n]

    def _extras(self, part, plan, idx, data_dict):
        if self.return_inverse:
            part["inverse"] = plan.inverse
        if self.return_grid_coord:
            part["grid_coord"] = ops.gather_rows(plan.grid_coord, idx)
            if "grid_coord" not in part["index_valid_keys"]:
                part["index_valid_keys"] = list(part["index_valid_keys"]) + ["grid_coord"]
        if self.return_min_coord:
            gs = np.asarray(plan.grid_size, dtype=np.float64)
            mc = np.asarray(plan.min_cell_host, dtype=np.int64) * gs          # transform.py:867
            part["min_coord"] = torch.from_numpy(mc.reshape(-1, 3)).to(idx.device)
        if self.return_displacement:
            disp = plan.displacement(idx)
            if self.project_displacement:
                disp = (disp * ops.gather_rows(data_dict["normal"], idx).to(disp.dtype)).sum(-1, keepdim=True)
            part["displacement"] = disp
            if "displacement" not in part["index_valid_keys"]:
                part["index_valid_keys"] = list(part["index_valid_keys"]) + ["displacement"]
        part["offset"] = torch.tensor(plan.new_offset_host, dtype=torch.int64, device=idx.device)

    def __call__(self, data_dict):
        assert "coord" in data_dict.keys()
        for k in ("sampled_index", "frame_pcd_offset"):
            if k in data_dict:
                raise NotImplementedError(f"GridSample: '{k}' is not supported by the batched CUDA transform")
        if "index_valid_keys" not in data_dict:
            data_dict["index_valid_keys"] = list(DEFAULT_INDEX_VALID_KEYS)
        plan = ops.grid_sample_plan(data_dict["coord"], self._offsets(data_dict), self.grid_size, self.hash_type, self.math)
        if self.mode == "train":
            seed = self.seed if self.seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
            idx = plan.select("train", seed)
            src = dict(data_dict)
            out = index_operator(dict(data_dict), idx)
            self._extras(out, plan, idx, src)
            return out
        parts = []
        for i in range(max(plan.count_max_host)):     # transform.py:913: one fragment per member rank
            idx = plan.select("test", i)
            part = index_operator(data_dict, idx, duplicate=True)
            part["index"] = idx
            self._extras(part, plan, idx, data_dict)
            parts.append(part)
        return parts


# ---------------------------------------------------------------------------------------------------------------------------
# on-disk scene layout (datasets/defaults.py:102-143): <root>/<split>/<scene>/{coord,color,normal,segment20,instance}.npy
# ---------------------------------------------------------------------------------------------------------------------------
def write_scene(path, coord, color=None, normal=None, segment=None, instance=None, segment_name="segment20"):
    os.makedirs(path, exist_ok=True)
    np.save(os.path.join(path, "coord.npy"), np.asarray(coord, dtype=np.float32))
    if color is not None:
        np.save(os.path.join(path, "color.npy"), np.asarray(color, dtype=np.uint8))
    if normal is not None:
        np.save(os.path.join(path, "normal.npy"), np.asarray(normal, dtype=np.float32))
    if segment is not None:
        np.save(os.path.join(path, f"{segment_name}.npy"), np.asarray(segment, dtype=np.int16))
    if instance is not None:
        np.save(os.path.join(path, "instance.npy"), np.asarray(instance, dtype=np.int16))


def load_scene(path, device="cuda", pin=True):
    """One scene directory -> dict of CUDA tensors with the dtypes DefaultDataset.get_data produces (defaults.py:118-143):
    coord / color / normal float32, segment / instance int32 (-1 filled when absent).  Host buffers are pinned so the copies
    are asynchronous."""
    out = {"name": os.path.basename(os.path.normpath(path)), "split": os.path.basename(os.path.dirname(os.path.normpath(path)))}
    for f in sorted(os.listdir(path)):
        if not f.endswith(".npy"):
            continue
        key = f[:-4]
        base = "segment" if key.startswith("segment") else key
        if base not in VALID_ASSETS:
            continue
        a = np.load(os.path.join(path, f))
        if base in ("coord", "color", "normal", "strength"):
            a = a.astype(np.float32)
        elif base in ("segment", "instance"):
            a = a.reshape(-1).astype(np.int32)
        t = torch.from_numpy(np.ascontiguousarray(a))
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        out[base] = t.to(device, non_blocking=True)
    n = out["coord"].shape[0]
    for k in ("segment", "instance"):
        if k not in out:
            out[k] = torch.full((n,), -1, dtype=torch.int32, device=out["coord"].device)
    return out
