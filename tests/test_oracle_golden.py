"""The oracle restatement vs fixtures produced by the reference's own code (tools/gen_golden.py)."""
import os

import numpy as np
import torch

from oracle import attention as oattn
from oracle import padding as opad
from oracle import serialization as oser


def test_encode_matches_reference_all_orders(golden_dir):
    g = np.load(os.path.join(golden_dir, "serialization.npz"))
    for name in ("d3", "d9", "d10", "d12", "d16"):
        depth = int(g[f"{name}_depth"])
        for order in oser.ORDERS:
            got = oser.encode(g[f"{name}_grid"], g[f"{name}_batch"], depth, order)
            assert np.array_equal(got, g[f"{name}_{order}"]), (name, order)


def test_encode_survey_appendix_b(golden_dir):
    g = np.load(os.path.join(golden_dir, "serialization.npz"))
    P = g["appB_points"]
    for depth in (3, 9, 16):
        for order in oser.ORDERS:
            assert np.array_equal(oser.encode(P, np.zeros(6, np.int64), depth, order), g[f"appB_d{depth}_{order}"])
    # SURVEY.md appendix B literal values
    assert oser.encode(P, np.zeros(6, np.int64), 3, "z").tolist() == [0, 4, 2, 1, 29, 511]
    assert oser.encode(P, np.zeros(6, np.int64), 3, "hilbert").tolist() == [0, 1, 7, 3, 48, 365]
    assert oser.encode(P, np.zeros(6, np.int64), 16, "hilbert").tolist() == [0, 7, 3, 1, 36, 365]
    assert oser.encode(P, np.zeros(6, np.int64), 16, "hilbert-trans").tolist() == [0, 3, 7, 1, 20, 365]


def test_survey_checksums():
    gen = torch.Generator().manual_seed(1234)
    G = torch.randint(0, 512, (100000, 3), generator=gen, dtype=torch.int32).numpy()
    B = torch.randint(0, 4, (100000,), generator=gen).sort().values.numpy()
    want = {"z": (26893477748927, 458716815), "z-trans": (26899619211059, 495942415),
            "hilbert": (26888520009227, 443170041), "hilbert-trans": (26899769319265, 511400043)}
    for order, (s, x) in want.items():
        code = oser.encode(G, B, 9, order)
        assert sum(int(c) for c in code) % (2 ** 61 - 1) == s
        assert int(np.bitwise_xor.reduce(code)) == x


def test_point_serialization(golden_dir):
    g = np.load(os.path.join(golden_dir, "point_padding.npz"))
    code, order, inverse, depth = oser.serialize(g["ser_grid"], g["ser_batch"], oser.ORDERS)
    assert depth == int(g["ser_depth"])
    assert np.array_equal(code, g["ser_code"])
    assert np.array_equal(order, g["ser_order"])
    assert np.array_equal(inverse, g["ser_inverse"])


def test_padding_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, "point_padding.npz"))
    for name in "abcdefgh":
        pad, unpad, cu = opad.padding_and_inverse(g[f"pad_{name}_offset"], int(g[f"pad_{name}_K"]))
        assert np.array_equal(pad, g[f"pad_{name}_pad"]), name
        assert np.array_equal(unpad, g[f"pad_{name}_unpad"]), name
        assert np.array_equal(cu, g[f"pad_{name}_cu"]), name
    pad, unpad, cu = opad.padding_and_inverse([5, 12], 4)
    assert pad.tolist() == [0, 1, 2, 3, 4, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 8]
    assert cu.tolist() == [0, 4, 8, 12, 16]


def test_dense_attention_matches_reference_branch(golden_dir):
    g = np.load(os.path.join(golden_dir, "attention_dense.npz"))
    H, C = int(g["H"]), int(g["C"])
    qkv_full = torch.from_numpy(g["qkv_full"])
    order = torch.from_numpy(g["order"])[torch.from_numpy(g["pad"])]
    inverse = torch.from_numpy(g["unpad"])[torch.from_numpy(g["inverse"])]
    qkv = qkv_full[order].reshape(-1, 3, H, C // H).clone().requires_grad_(True)
    out = oattn.varlen_attention(qkv, g["cu"], float(g["scale"]))
    core = out.reshape(-1, C)[inverse]
    assert torch.allclose(core, torch.from_numpy(g["core_out"]), atol=2e-6, rtol=1e-5)
    core.backward(torch.from_numpy(g["d_core_out"]))
    d_full = torch.zeros_like(qkv_full).index_add_(0, order, qkv.grad.reshape(-1, 3 * C))
    assert torch.allclose(d_full, torch.from_numpy(g["d_qkv_full"]), atol=2e-6, rtol=1e-4)


# ---- GridSample (pointcept/datasets/transform.py:840-958) -----------------------------------------------------------------
def test_grid_sample_oracle_matches_reference_fixture(golden_dir):
    from oracle import grid_sample as ogs
    g = np.load(os.path.join(golden_dir, "grid_sample.npz"))
    assert np.array_equal(ogs.fnv_hash_vec(np.arange(30, dtype=np.int64).reshape(10, 3)), g["fnv_of_arange"])
    for hash_type in ("fnv", "ravel"):
        for gs in (0.05, 0.02):
            for i in range(2):
                tag = f"{hash_type}_{gs}_{i}"
                p = ogs.plan(g[f"coord{i}"], gs, hash_type, "float64")
                assert np.array_equal(p["inverse"], g[tag + "_inverse"]), tag
                assert int(p["count"].max()) == int(g[tag + "_n_fragments"]), tag
                assert np.array_equal(p["grid_coord"][ogs.select(p, 0)], g[tag + "_grid_coord"]), tag
                assert np.array_equal((p["min_cell"] * np.float64(gs)).reshape(1, 3), g[tag + "_min_coord"]), tag
                # whichever member the reference's (unstable) argsort put first, it lies in the voxel the oracle assigns
                for key in ("_index0", "_last_index"):
                    assert np.array_equal(p["inverse"][g[tag + key]], np.arange(len(p["count"]))), tag
                if hash_type == "fnv":
                    disp = (p["scaled"] - p["grid_coord"] - 0.5)[g[tag + "_index0"]]
                    assert np.array_equal(disp, g[tag + "_displacement0"]), tag


# ---- PointROPE (pointcept/models/litept/litept_v1.py:66-125) -------------------------------------------------------------
def test_point_rope_oracle_matches_reference_fixture(golden_dir):
    from oracle import eval_ops as oev
    g = np.load(os.path.join(golden_dir, "point_rope.npz"))
    for name in ("d18", "d48"):
        y = oev.point_rope(torch.from_numpy(g[name + "_tokens"]), torch.from_numpy(g[name + "_pos"]), float(g[name + "_base"]))
        assert float((y - torch.from_numpy(g[name + "_out"])).abs().max()) <= 1e-6, name
        # rotation by the negative angle is the inverse (what PointROPE_func.backward applies, litept_v1.py:40-46)
        back = oev.point_rope(y, torch.from_numpy(g[name + "_pos"]), float(g[name + "_base"]), -1.0)
        assert float((back - torch.from_numpy(g[name + "_tokens"])).abs().max()) <= 1e-5


def test_knn_oracle_corroborated_by_scipy_ckdtree():
    """The reference's knn_query is CUDA-only (libs/pointops/src/knn_query) and ships no vectors, so oracle/eval_ops.knn_query cannot be
    pinned; it is checked here against an independent exact k-NN (scipy's cKDTree): per-scene search, ascending order, global row
    indices, Euclidean distance, and the -1 / 1e5 placeholders when a scene holds fewer than k points."""
    import numpy as np
    from scipy.spatial import cKDTree
    from oracle import eval_ops
    rng = np.random.default_rng(5)
    sizes, qsizes, k = [300, 7, 120], [40, 5, 33], 9
    xyz = rng.uniform(-2, 2, (sum(sizes), 3)).astype(np.float32)
    new = rng.uniform(-2, 2, (sum(qsizes), 3)).astype(np.float32)
    offset, new_offset = np.cumsum(sizes), np.cumsum(qsizes)
    idx, dist = eval_ops.knn_query(k, xyz, offset, new, new_offset)
    s = qs = 0
    for e, qe in zip(offset, new_offset):
        kk = min(k, e - s)
        d_ref, i_ref = cKDTree(xyz[s:e].astype(np.float64)).query(new[qs:qe].astype(np.float64), k=kk)
        d_ref, i_ref = d_ref.reshape(qe - qs, kk), i_ref.reshape(qe - qs, kk)
        assert np.array_equal(idx[qs:qe, :kk], i_ref + s)                    # same neighbours, same (ascending) order, global rows
        assert np.allclose(dist[qs:qe, :kk], d_ref, rtol=0, atol=2e-6)
        assert (idx[qs:qe, kk:] == -1).all() and np.allclose(dist[qs:qe, kk:], 1e5)      # sqrt(1e10) placeholders
        s, qs = e, qe
