"""Load the UNMODIFIED reference python files from /root/reference in this container.

Only used by tools/gen_golden.py (fixture generation) and by CPU tests that skip
when /root/reference is absent (it does not exist on the GPU box).  Third-party
modules the reference imports but this image lacks are replaced by minimal stand-ins
for names that are NOT on the hot path (addict.Dict, timm DropPath, torch_scatter,
HookBase ...); ``spconv`` / ``flash_attn`` can be either stubs or our drop-in shims.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("POINTCEPT_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "pointcept"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


class _AttrDict(dict):
    """stand-in for addict.Dict (attribute access on a dict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def load_serialization():
    """-> module with encode(), z_order, hilbert exactly as the reference defines them."""
    pkg = "_ref_serialization"
    p = _mod(pkg)
    p.__path__ = [os.path.join(REF, "pointcept/models/utils/serialization")]
    _load(pkg + ".z_order", "pointcept/models/utils/serialization/z_order.py")
    _load(pkg + ".hilbert", "pointcept/models/utils/serialization/hilbert.py")
    return _load(pkg + ".default", "pointcept/models/utils/serialization/default.py")


def _install_oracle_spconv():
    """CPU stand-in for spconv built on oracle/spconv_ref.py, so the UNMODIFIED reference model files run
    on CPU (fixture generation / pinning the model-level restatement; spconv arithmetic itself stays unpinned)."""
    import numpy as np
    import torch
    import torch.nn as nn
    from oracle import spconv_ref as osp

    class SparseConvTensor:
        def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None, **kw):
            self.features, self.indices = features, indices
            self.spatial_shape, self.batch_size = list(spatial_shape), batch_size
            self.indice_dict = indice_dict if indice_dict is not None else {}

        def replace_feature(self, f):
            return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)

    class SparseModule(nn.Module):
        pass

    class SubMConv3d(SparseModule):
        def __init__(self, cin, cout, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, indice_key=None, **kw):
            super().__init__()
            self.k, self.indice_key = kernel_size, indice_key
            self.weight = nn.Parameter(torch.empty(cout, kernel_size, kernel_size, kernel_size, cin))
            self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
            nn.init.normal_(self.weight, std=0.05)

        def forward(self, x):
            key = (self.indice_key, self.k)
            if self.indice_key is None or key not in x.indice_dict:
                pair = osp.subm_rulebook(x.indices.numpy(), x.spatial_shape, self.k)
                if self.indice_key is not None:
                    x.indice_dict[key] = pair
            else:
                pair = x.indice_dict[key]
            w = self.weight.reshape(self.weight.shape[0], -1, self.weight.shape[-1])
            return x.replace_feature(osp.conv_apply(x.features, w, pair, self.bias))

    class SparseConv3d(SparseModule):
        """strided sparse conv on the oracle rulebook; remembers, under its indice_key, what SparseInverseConv3d needs"""
        def __init__(self, cin, cout, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, indice_key=None, **kw):
            super().__init__()
            self.k, self.s, self.p, self.indice_key = kernel_size, stride, padding, indice_key
            self.weight = nn.Parameter(torch.empty(cout, kernel_size, kernel_size, kernel_size, cin))
            self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
            nn.init.normal_(self.weight, std=0.05)

        def forward(self, x):
            oi, osh, pf, pb = osp.strided_rulebook(x.indices.numpy(), x.spatial_shape, self.k, self.s, self.p)
            if self.indice_key is not None:
                x.indice_dict[("strided", self.indice_key)] = (x.indices, x.spatial_shape, pb)
            w = self.weight.reshape(self.weight.shape[0], -1, self.weight.shape[-1])
            return SparseConvTensor(osp.conv_apply(x.features, w, pf, self.bias), torch.from_numpy(oi), osh, x.batch_size, x.indice_dict)

    class SparseInverseConv3d(SparseModule):
        def __init__(self, cin, cout, kernel_size, indice_key=None, bias=True, **kw):
            super().__init__()
            self.k, self.indice_key = kernel_size, indice_key
            self.weight = nn.Parameter(torch.empty(cout, kernel_size, kernel_size, kernel_size, cin))
            self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
            nn.init.normal_(self.weight, std=0.05)

        def forward(self, x):
            indices, shape, pb = x.indice_dict[("strided", self.indice_key)]
            w = self.weight.reshape(self.weight.shape[0], -1, self.weight.shape[-1])
            return SparseConvTensor(osp.inverse_conv_apply(x.features, w, pb, self.bias), indices, shape, x.batch_size, x.indice_dict)

    class SparseSequential(SparseModule):
        """spconv's container: sparse modules take the tensor, plain modules (BatchNorm1d, ReLU, Identity) its features"""
        def __init__(self, *args, **kwargs):
            super().__init__()
            import collections
            if len(args) == 1 and isinstance(args[0], collections.OrderedDict):
                for k, m in args[0].items():
                    self.add_module(k, m)
            else:
                for i, m in enumerate(args):
                    self.add_module(str(i), m)
            for k, m in kwargs.items():
                self.add_module(k, m)

        def __getitem__(self, i):
            return list(self._modules.values())[i]

        def __len__(self):
            return len(self._modules)

        def forward(self, x):
            for m in self._modules.values():
                if isinstance(m, SparseModule):
                    x = m(x)
                elif isinstance(x, SparseConvTensor):
                    x = x.replace_feature(m(x.features))
                else:
                    x = m(x)
            return x

    sp = _mod("spconv")
    spt = _mod("spconv.pytorch", SubMConv3d=SubMConv3d, SparseConv3d=SparseConv3d, SparseInverseConv3d=SparseInverseConv3d,
               SparseModule=SparseModule, SparseSequential=SparseSequential, Identity=nn.Identity,
               SparseConvTensor=SparseConvTensor)
    spt.modules = _mod("spconv.pytorch.modules", is_spconv_module=lambda m: isinstance(m, SparseModule))
    sp.pytorch = spt


def load_models(use_shims=False):
    """Import structure.py, modules.py, PT-v3m1 and SpUNet-v1m1 from the reference.

    use_shims=True wires ``spconv`` / ``flash_attn`` to pointcept_b200's drop-in modules
    (the boundary test); otherwise inert stubs are used (CPU oracle / golden generation).
    """
    import torch
    import torch.nn as nn

    _mod("addict", Dict=_AttrDict)

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    _mod("timm")
    _mod("timm.layers", DropPath=DropPath, trunc_normal_=nn.init.trunc_normal_)

    def segment_csr(src, indptr, reduce="sum"):
        out = []
        for a, b in zip(indptr[:-1].tolist(), indptr[1:].tolist()):
            seg = src[a:b]
            out.append({"sum": seg.sum(0), "mean": seg.mean(0), "max": seg.max(0).values,
                        "min": seg.min(0).values}[reduce])
        return torch.stack(out)

    _mod("torch_scatter", segment_csr=segment_csr)
    _mod("torch_geometric")
    _mod("torch_geometric.utils", scatter=None)

    if use_shims == "oracle":
        _install_oracle_spconv()
        sys.modules["flash_attn"] = None
    elif use_shims:
        import pointcept_b200
        pointcept_b200.install(flash_attn=True)
    else:
        class _Stub(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
        sp = _mod("spconv")
        spt = _mod("spconv.pytorch", SubMConv3d=_Stub, SparseConv3d=_Stub, SparseInverseConv3d=_Stub,
                   SparseModule=nn.Module, SparseSequential=nn.Sequential, Identity=nn.Identity,
                   SparseConvTensor=object)
        spt.modules = _mod("spconv.pytorch.modules", is_spconv_module=lambda m: isinstance(m, _Stub))
        sp.pytorch = spt
        sys.modules["flash_attn"] = None  # "import flash_attn" raises ImportError -> reference sets it to None

    # parent packages (empty shells so relative/absolute imports of the files below resolve)
    for name in ["pointcept", "pointcept.models", "pointcept.models.utils", "pointcept.engines",
                 "pointcept.models.point_prompt_training", "pointcept.utils"]:
        if name not in sys.modules:
            m = _mod(name)
            m.__path__ = []

    class HookBase:
        pass

    _mod("pointcept.engines.hooks", HookBase=HookBase)

    class _Registry:
        def register_module(self, name=None, force=False, module=None):
            def deco(cls):
                return cls
            return deco

    _mod("pointcept.models.builder", MODELS=_Registry(), MODULES=_Registry())
    _mod("pointcept.models.point_prompt_training", PDNorm=None)
    ser = "pointcept.models.utils.serialization"
    sp_ = _mod(ser)
    sp_.__path__ = [os.path.join(REF, "pointcept/models/utils/serialization")]
    _load(ser + ".z_order", "pointcept/models/utils/serialization/z_order.py")
    _load(ser + ".hilbert", "pointcept/models/utils/serialization/hilbert.py")
    d = _load(ser + ".default", "pointcept/models/utils/serialization/default.py")
    sp_.encode = d.encode
    misc = _load("pointcept.models.utils.misc", "pointcept/models/utils/misc.py")
    u = sys.modules["pointcept.models.utils"]
    for k in ("offset2batch", "batch2offset", "offset2bincount", "bincount2offset"):
        setattr(u, k, getattr(misc, k))
    structure = _load("pointcept.models.utils.structure", "pointcept/models/utils/structure.py")
    modules = _load("pointcept.models.modules", "pointcept/models/modules.py")
    ptv3 = _load("pointcept.models.point_transformer_v3.point_transformer_v3m1_base",
                 "pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py")
    spunet = _load("pointcept.models.sparse_unet.spconv_unet_v1m1_base",
                   "pointcept/models/sparse_unet/spconv_unet_v1m1_base.py")
    return types.SimpleNamespace(structure=structure, modules=modules, ptv3=ptv3, spunet=spunet, misc=misc)


def load_transform():
    """-> the reference's pointcept/datasets/transform.py (GridSample, index_operator ...) under a private package name, with
    only its Registry dependency loaded from the reference tree."""
    pkg = "_ref_datasets"
    if pkg + ".transform" in sys.modules:
        return sys.modules[pkg + ".transform"]
    saved = {k: sys.modules.get(k) for k in ("pointcept", "pointcept.utils", "pointcept.utils.misc", "pointcept.utils.registry")}
    try:
        p = _mod("pointcept")
        p.__path__ = [os.path.join(REF, "pointcept")]
        u = _mod("pointcept.utils")
        u.__path__ = [os.path.join(REF, "pointcept/utils")]
        _load("pointcept.utils.misc", "pointcept/utils/misc.py")
        _load("pointcept.utils.registry", "pointcept/utils/registry.py")
        return _load(pkg + ".transform", "pointcept/datasets/transform.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_litept():
    """-> the reference's pointcept/models/litept/litept_v1.py.  ``pointrope`` (its CUDA extension) is not importable here, so the
    module defines its pure-PyTorch PointROPE fallback (litept_v1.py:66-125) -- the CPU reference for the rotary embedding."""
    load_models(use_shims=False)
    sys.modules.pop("pointrope", None)
    if "pointcept.models.litept" not in sys.modules:
        m = _mod("pointcept.models.litept")
        m.__path__ = []
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return _load("pointcept.models.litept.litept_v1", "pointcept/models/litept/litept_v1.py")
