"""SubMConv3d / SparseConv3d / SparseInverseConv3d with spconv's constructor signature, parameter names
and shapes (weight [Cout, K0, K1, K2, Cin], bias [Cout]) so checkpoints and optimizer param groups of the
reference load unchanged (pointcept/engines/hooks/misc.py:236-274, pointcept/utils/optimizer.py:43-52)."""
import math

import torch
import torch.nn as nn

from ... import ops
from .core import IndiceData, SparseConvTensor
from .modules import SparseModule


def _triple(v, ndim=3):
    if isinstance(v, (list, tuple)):
        assert len(v) == ndim
        return [int(x) for x in v]
    return [int(v)] * ndim


class SparseConvolution(SparseModule):
    _b2pc_half_shadow = True     # ops.HalfShadows keeps half-precision copies of weight / bias for the autocast path

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 algo=None, fp32_accum=None, record_voxel_count=False, act_type=None, act_alpha=0, act_beta=0,
                 large_kernel_fast_algo=False, name=None):
        super().__init__()
        if ndim != 3:
            raise NotImplementedError("pointcept_b200.spconv implements 3D convolutions only")
        if groups != 1:
            raise NotImplementedError("groups != 1 is not supported")
        if transposed:
            raise NotImplementedError("SparseConvTranspose3d is not on the Pointcept hot path")
        self.ndim = ndim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        self.output_padding = _triple(output_padding)
        if subm and any(k % 2 == 0 for k in self.kernel_size):
            # the backward-data pass reads the forward rulebook with flipped offsets, valid only for symmetric (odd) kernels
            raise NotImplementedError("SubMConv3d with an even kernel size is not supported (no stock Pointcept config uses one)")
        self.conv1x1 = all(k == 1 for k in self.kernel_size)
        self.subm, self.inverse, self.transposed = subm, inverse, transposed
        self.groups = groups
        self.indice_key = indice_key
        self.algo = algo
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def extra_repr(self):
        s = "{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}"
        if any(p != 0 for p in self.padding):
            s += ", padding={padding}"
        if self.bias is None:
            s += ", bias=False"
        s += ", subm={subm}, inverse={inverse}, indice_key={indice_key}"
        return s.format(**self.__dict__)

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    # ---- rulebook ---------------------------------------------------------------------------------
    def _rulebook(self, x):
        key = self.indice_key
        data = x.find_indice_pair(key)
        if self.inverse:
            if data is None:
                raise RuntimeError(f"SparseInverseConv3d: no rulebook stored under indice_key={key!r}")
            if data.is_subm or data.ksize != self.kernel_size:
                raise RuntimeError("SparseInverseConv3d: kernel geometry differs from the paired SparseConv3d")
            return data
        if data is not None:
            if data.ksize != self.kernel_size or data.is_subm != self.subm or (
                    not self.subm and (data.stride != self.stride or data.padding != self.padding)):
                raise RuntimeError(f"indice_key={key!r} is already used by a convolution with different geometry")
            if data.indices.shape[0] != x.indices.shape[0]:
                raise RuntimeError(f"indice_key={key!r}: rulebook was built for a different voxel set")
            return data
        if self.subm:
            pair = ops.rulebook_subm(x.indices, x.spatial_shape, self.kernel_size, self.dilation)
            data = IndiceData(subm=True, ksize=self.kernel_size, stride=[1, 1, 1], padding=self.padding,
                              dilation=self.dilation, in_indices=x.indices, out_indices=x.indices,
                              in_shape=x.spatial_shape, out_shape=x.spatial_shape, pair_fwd=pair, pair_bwd=pair)
        else:
            out_idx, out_shape, pf, pb = ops.rulebook_strided(x.indices, x.spatial_shape, self.kernel_size, self.stride,
                                                              self.padding, self.dilation)
            data = IndiceData(subm=False, ksize=self.kernel_size, stride=self.stride, padding=self.padding,
                              dilation=self.dilation, in_indices=x.indices, out_indices=out_idx,
                              in_shape=x.spatial_shape, out_shape=out_shape, pair_fwd=pf, pair_bwd=pb)
        if key is not None:
            x.indice_dict[key] = data
        return data

    def forward(self, x):
        assert isinstance(x, SparseConvTensor)
        feat = x.features
        w16 = b16 = None
        if torch.is_autocast_enabled():
            adt = torch.get_autocast_dtype("cuda")
            half = getattr(x, "_features_half", None)     # emitted by the producer of x.features (fused residual kernel)
            feat = half if (half is not None and half.dtype == adt and half.shape == feat.shape) else feat.to(adt)
            w16, b16 = ops.shadow_of(self, adt)
        kv = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        w = self.weight.view(self.out_channels, kv, self.in_channels)
        if self.conv1x1 and self.subm:
            # K = 1: plain features @ W^T + b (library GEMM); no rulebook needed
            if w16 is not None:
                out = ops.linear(feat, w[:, 0, :], self.bias, w16.view(self.out_channels, self.in_channels), b16)
            else:
                out = torch.nn.functional.linear(feat, w[:, 0, :].to(feat.dtype),
                                                 self.bias.to(feat.dtype) if self.bias is not None else None)
            return x.replace_feature(out)
        data = self._rulebook(x)
        if self.in_channels % 16 != 0 and feat.dtype in (torch.float16, torch.bfloat16):
            # tensor-core tiles want the reduction axis in multiples of 16 channels (e.g. the 6-channel stem): zero-pad
            # features and weight; autograd slices the gradients back
            padc = 16 - self.in_channels % 16
            feat = torch.nn.functional.pad(feat, (0, padc))
            w = torch.nn.functional.pad(w, (0, padc))
            w16 = b16 = None
        if self.inverse:
            out = ops.sparse_conv(feat, w, self.bias, data.pair_bwd, data.pair_fwd, False, w16, b16)
            res = SparseConvTensor(out, data.indices, data.spatial_shape, x.batch_size, x.grid, x.voxel_num,
                                   x.indice_dict, x.benchmark)
        elif self.subm:
            out = ops.sparse_conv(feat, w, self.bias, data.pair_fwd, data.pair_fwd, True, w16, b16)
            res = x.replace_feature(out)
        else:
            out = ops.sparse_conv(feat, w, self.bias, data.pair_fwd, data.pair_bwd, False, w16, b16)
            res = SparseConvTensor(out, data.out_indices, data.out_spatial_shape, x.batch_size, x.grid, x.voxel_num,
                                   x.indice_dict, x.benchmark)
        res.benchmark_record = x.benchmark_record
        return res


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, large_kernel_fast_algo=False, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key, algo=algo, fp32_accum=fp32_accum,
                         large_kernel_fast_algo=large_kernel_fast_algo, name=name)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, record_voxel_count=False, large_kernel_fast_algo=False,
                 name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key, algo=algo, fp32_accum=fp32_accum, record_voxel_count=record_voxel_count,
                         large_kernel_fast_algo=large_kernel_fast_algo, name=name)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, algo=None, fp32_accum=None,
                 large_kernel_fast_algo=False, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key,
                         algo=algo, fp32_accum=fp32_accum, large_kernel_fast_algo=large_kernel_fast_algo, name=name)
