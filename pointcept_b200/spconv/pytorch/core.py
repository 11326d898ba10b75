import torch


class IndiceData:
    """Rulebook cached under an ``indice_key`` (the role spconv's ImplicitGemmIndiceData plays)."""

    def __init__(self, *, subm, ksize, stride, padding, dilation, in_indices, out_indices, in_shape, out_shape,
                 pair_fwd, pair_bwd):
        self.is_subm = subm
        self.ksize, self.stride, self.padding, self.dilation = ksize, stride, padding, dilation
        self.indices = in_indices            # [N_in, 4]
        self.out_indices = out_indices       # [N_out, 4]
        self.spatial_shape = in_shape
        self.out_spatial_shape = out_shape
        self.pair_fwd = pair_fwd             # [KV, N_out] input row per output row
        self.pair_bwd = pair_bwd             # [KV, N_in] output row per input row (subm: pair_fwd, read flipped)


class SparseConvTensor:
    """features [N,C], indices [N,4] int32 (batch, x, y, z), spatial_shape, batch_size, shared indice_dict."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False, permanent_thrust_allocator=False, enable_timer=False, force_algo=None):
        assert features.dim() == 2 and indices.dim() == 2 and indices.shape[1] == 4
        assert indices.dtype == torch.int32, "indices must be int32 (batch, x, y, z)"
        assert features.shape[0] == indices.shape[0]
        self._features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self.benchmark_record = {}
        self.thrust_allocator = None
        self.force_algo = force_algo
        self.int8_scale = None

    @property
    def features(self):
        return self._features

    @features.setter
    def features(self, val):
        raise ValueError("use x = x.replace_feature(feat) (spconv >= 2.1 semantics)")

    def replace_feature(self, feature):
        t = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.voxel_num,
                             self.indice_dict, self.benchmark)
        t.benchmark_record = self.benchmark_record
        return t

    def shadow_copy(self):
        return self.replace_feature(self._features)

    @property
    def spatial_size(self):
        r = 1
        for s in self.spatial_shape:
            r *= s
        return r

    @property
    def is_quantized(self):
        return False

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def dense(self, channels_first=True):
        b, shp = self.batch_size, self.spatial_shape
        out = self._features.new_zeros((b, *shp, self._features.shape[1]))
        idx = self.indices.long()
        out[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = self._features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out
