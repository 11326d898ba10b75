"""``spconv.pytorch`` drop-in: the names Pointcept's PT-v3m1 / SpUNet-v1m1 use, same signatures.

Reference call sites: pointcept/models/utils/structure.py:139-146 (SparseConvTensor),
point_transformer_v3m1_base.py:278-284,499-506 (SubMConv3d), sparse_unet/spconv_unet_v1m1_base.py:23,
40-47,113-124,136-147,172-182,222-226 (SparseModule, SparseSequential, SubMConv3d, SparseConv3d,
SparseInverseConv3d, Identity), models/modules.py:84 (modules.is_spconv_module).
"""
from . import modules  # noqa: F401
from .conv import SparseConv3d, SparseConvolution, SparseInverseConv3d, SubMConv3d  # noqa: F401
from .core import IndiceData, SparseConvTensor  # noqa: F401
from .modules import Identity, SparseModule, SparseSequential  # noqa: F401
