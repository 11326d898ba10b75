"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/b2pc.h declares,
and validates arguments before touching the GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

from pointcept_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "b2pc.h")).read()
    declared = set(re.findall(r"\b(b2pc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in b2pc.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.b2pc_version() >= 100


def test_argument_errors_are_reported_without_a_gpu(lib):
    fake = ctypes.c_void_p(0x1000)
    orders = (ctypes.c_int * 1)(0)
    rc = lib.b2pc_serialize_encode(fake, None, 10, 99, orders, 1, fake, None)
    assert rc == -1 and b"depth" in lib.b2pc_last_error()
    rc = lib.b2pc_serialize_encode(None, None, 10, 9, orders, 1, fake, None)
    assert rc == -1 and b"null" in lib.b2pc_last_error()
    rc = lib.b2pc_patch_attn_fwd(fake, 0, fake, 1, 16, 16, 1, 16, 0.25, fake, fake, 0, None)
    assert rc == -1 and b"dtype" in lib.b2pc_last_error()
    rc = lib.b2pc_serialize_sort(fake, 1000, 4, 30, fake, fake, fake, 16, None)
    assert rc == -2 and b"workspace" in lib.b2pc_last_error()
    with pytest.raises(RuntimeError, match="workspace"):
        _lib.check(rc, "serialize_sort")


def test_workspace_queries(lib):
    assert lib.b2pc_serialize_sort_workspace_bytes(100000, 4) >= 100000 * 4 * (8 + 4) * 2
    assert lib.b2pc_rulebook_workspace_bytes(100000, 1) > 100000 * 12
    assert lib.b2pc_spconv_bwd_weight_workspace_bytes(120000, 32, 32, 27) >= 27 * 32 * 32 * 4
    assert lib.b2pc_patch_attn_bwd_workspace_bytes(1 << 17, 2, 16) >= (1 << 17) * 2 * 4


def test_ops_refuse_cpu_tensors():
    import torch
    from pointcept_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.serialize_encode(torch.zeros(4, 3, dtype=torch.int32), torch.zeros(4, dtype=torch.long), 4, ["z"])
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.patch_attention(torch.zeros(8, 3, 1, 16, dtype=torch.bfloat16), torch.tensor([0, 8], dtype=torch.int32), 8)
