// Hash-grid over active voxels and rulebook (indice-pair) generation.
// Rulebook form: pair[KV, N_out] int32 (input row or -1).  Integer / HBM-L2 bound:
// 16 B/voxel in, 4*KV B/voxel out; the hash table (<= 24 B/voxel) stays L2 resident.
#pragma once
#include "common.cuh"
#include "sort.cuh"

namespace b2pc {

constexpr uint64_t kEmptyKey = ~0ull;

struct Geometry {
  int shape[3];    // input spatial shape
  int oshape[3];   // output spatial shape (strided)
  int k[3], s[3], p[3], d[3];
  int kv;
};

__device__ __forceinline__ uint32_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}

__device__ __forceinline__ uint64_t lin_key(int b, int x, int y, int z, const int* shape) {
  return (((uint64_t)b * shape[0] + x) * shape[1] + y) * shape[2] + z;
}

__global__ void __launch_bounds__(256)
hash_clear_kernel(uint64_t* __restrict__ keys, int64_t slots) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < slots; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = kEmptyKey;
}

// insert key -> row.  Duplicate coordinates keep the smallest row (deterministic).
__global__ void __launch_bounds__(256)
hash_insert_rows_kernel(const int4* __restrict__ indices, int64_t n, Geometry g, uint64_t* __restrict__ keys,
                        int32_t* __restrict__ vals, uint32_t mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = indices[i];
    const uint64_t key = lin_key(c.x, c.y, c.z, c.w, g.shape);
    uint32_t slot = hash64(key) & mask;
    while (true) {
      const uint64_t prev = atomicCAS((unsigned long long*)&keys[slot], (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (prev == kEmptyKey || prev == key) { atomicMin(&vals[slot], (int32_t)i); break; }
      slot = (slot + 1) & mask;
    }
  }
}

__device__ __forceinline__ int32_t hash_lookup(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                               uint32_t mask, uint64_t key) {
  uint32_t slot = hash64(key) & mask;
  while (true) {
    const uint64_t k = keys[slot];
    if (k == key) return vals[slot];
    if (k == kEmptyKey) return -1;
    slot = (slot + 1) & mask;
  }
}

// one thread per voxel, loop over kernel offsets: coordinates read once, every pair[k, :] write is
// warp-coalesced.
__global__ void __launch_bounds__(256)
subm_pairs_kernel(const int4* __restrict__ indices, int64_t n, Geometry g, const uint64_t* __restrict__ keys,
                  const int32_t* __restrict__ vals, uint32_t mask, int32_t* __restrict__ pair) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = indices[j];
    int k = 0;
    for (int i0 = 0; i0 < g.k[0]; ++i0) {
      const int x = c.y + (i0 - g.k[0] / 2) * g.d[0];
      for (int i1 = 0; i1 < g.k[1]; ++i1) {
        const int y = c.z + (i1 - g.k[1] / 2) * g.d[1];
#pragma unroll 4
        for (int i2 = 0; i2 < g.k[2]; ++i2, ++k) {
          const int z = c.w + (i2 - g.k[2] / 2) * g.d[2];
          int32_t r = -1;
          if (x >= 0 && x < g.shape[0] && y >= 0 && y < g.shape[1] && z >= 0 && z < g.shape[2])
            r = hash_lookup(keys, vals, mask, lin_key(c.x, x, y, z, g.shape));
          pair[(int64_t)k * n + j] = r;
        }
      }
    }
  }
}

// ---- strided ------------------------------------------------------------------------------------
// out key of input voxel c through offset (i0,i1,i2), or kEmptyKey
__device__ __forceinline__ uint64_t strided_out_key(const int4 c, int i0, int i1, int i2, const Geometry& g) {
  const int cc[3] = {c.y, c.z, c.w};
  const int ii[3] = {i0, i1, i2};
  int o[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int num = cc[a] + g.p[a] - ii[a] * g.d[a];
    if (num < 0 || num % g.s[a] != 0) return kEmptyKey;
    o[a] = num / g.s[a];
    if (o[a] >= g.oshape[a]) return kEmptyKey;
  }
  return lin_key(c.x, o[0], o[1], o[2], g.oshape);
}

// phase 1: insert every reachable out key into the set; first inserter appends it to the list.
__global__ void __launch_bounds__(256)
strided_collect_kernel(const int4* __restrict__ indices, int64_t n, Geometry g, uint64_t* __restrict__ keys,
                       uint32_t mask, uint64_t* __restrict__ uniq, unsigned long long* __restrict__ num_out) {
  int bmax = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bmax = max(bmax, indices[i].x);
  bmax = __reduce_max_sync(0xFFFFFFFFu, bmax);
  if ((threadIdx.x & 31) == 0 && bmax > 0) atomicMax(num_out + 1, (unsigned long long)bmax);   // num_out[1] = largest batch index
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = indices[i];
    for (int i0 = 0; i0 < g.k[0]; ++i0)
      for (int i1 = 0; i1 < g.k[1]; ++i1)
        for (int i2 = 0; i2 < g.k[2]; ++i2) {
          const uint64_t key = strided_out_key(c, i0, i1, i2, g);
          if (key == kEmptyKey) continue;
          uint32_t slot = hash64(key) & mask;
          while (true) {
            const uint64_t prev = atomicCAS((unsigned long long*)&keys[slot], (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (prev == kEmptyKey) { uniq[atomicAdd(num_out, 1ull)] = key; break; }
            if (prev == key) break;
            slot = (slot + 1) & mask;
          }
        }
  }
}

// phase 2 (after the distinct keys were sorted): rank -> table value, out_indices
__global__ void __launch_bounds__(256)
strided_assign_kernel(const int64_t* __restrict__ order, const uint64_t* __restrict__ uniq, int64_t m, Geometry g,
                      const uint64_t* __restrict__ keys, int32_t* __restrict__ vals, uint32_t mask,
                      int32_t* __restrict__ out_indices) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = uniq[order[r]];
    uint32_t slot = hash64(key) & mask;
    while (keys[slot] != key) slot = (slot + 1) & mask;
    vals[slot] = (int32_t)r;
    uint64_t rem = key;
    const int z = (int)(rem % g.oshape[2]); rem /= g.oshape[2];
    const int y = (int)(rem % g.oshape[1]); rem /= g.oshape[1];
    const int x = (int)(rem % g.oshape[0]); rem /= g.oshape[0];
    reinterpret_cast<int4*>(out_indices)[r] = make_int4((int)rem, x, y, z);
  }
}

__global__ void __launch_bounds__(256)
fill_i32_kernel(int32_t* __restrict__ p, int64_t n, int32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// phase 3: both direction tables.  pair_bwd[k, i] = out row; pair_fwd[k, out row] = i
// (for a fixed k an out row has at most one source, so the write is race free).
__global__ void __launch_bounds__(256)
strided_pairs_kernel(const int4* __restrict__ indices, int64_t n, int64_t m, Geometry g, const uint64_t* __restrict__ keys,
                     const int32_t* __restrict__ vals, uint32_t mask, int32_t* __restrict__ pair_fwd,
                     int32_t* __restrict__ pair_bwd) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = indices[i];
    int k = 0;
    for (int i0 = 0; i0 < g.k[0]; ++i0)
      for (int i1 = 0; i1 < g.k[1]; ++i1)
        for (int i2 = 0; i2 < g.k[2]; ++i2, ++k) {
          const uint64_t key = strided_out_key(c, i0, i1, i2, g);
          int32_t r = -1;
          if (key != kEmptyKey) {
            r = hash_lookup(keys, vals, mask, key);
            pair_fwd[(int64_t)k * m + r] = (int32_t)i;
          }
          pair_bwd[(int64_t)k * n + i] = r;
        }
  }
}

// ---- host side ------------------------------------------------------------------------------------
inline int64_t hash_slots(int64_t n_keys) {
  int64_t s = 1024;
  while (s < 2 * n_keys) s <<= 1;
  return s;
}

struct RulebookWs {
  uint64_t* keys; int32_t* vals; uint64_t* uniq; int64_t* order; int64_t* inverse; void* sort_ws;
  size_t sort_bytes; int64_t slots; size_t total;
};

// reach = how many distinct outputs one input voxel can feed: per axis the offsets i with (c + p - i*d) % s == 0, i.e.
// ceil(k / (s / gcd(s, d))) of them (1 for the k = 2, s = 2 down-convolutions of SpUNet, 8 for k = 3, s = 2)
inline int strided_reach(const int* k, const int* s, const int* d) {
  int reach = 1;
  for (int a = 0; a < 3; ++a) {
    int sa = s ? s[a] : 1, da = d ? d[a] : 1, x = sa, y = da;
    while (y) { const int t = x % y; x = y; y = t; }
    const int period = sa / (x > 0 ? x : 1);
    reach *= (k[a] + period - 1) / period;
  }
  return reach < 1 ? 1 : reach;
}

inline RulebookWs carve_rulebook_ws(void* ws, int64_t n, int reach) {
  RulebookWs r;
  const int64_t cap = n * (int64_t)(reach < 1 ? 1 : reach);  // most distinct keys a strided build can see
  r.slots = hash_slots(cap > n ? cap : n);
  char* p = (char*)ws;
  size_t off = 0;
  r.keys = (uint64_t*)(p + off); off += align_up((size_t)r.slots * 8, 256);
  r.vals = (int32_t*)(p + off); off += align_up((size_t)r.slots * 4, 256);
  r.uniq = (uint64_t*)(p + off); off += align_up((size_t)cap * 8, 256);
  r.order = (int64_t*)(p + off); off += align_up((size_t)cap * 8, 256);
  r.inverse = (int64_t*)(p + off); off += align_up((size_t)cap * 8, 256);
  r.sort_bytes = sort_workspace_bytes(cap, 1);
  r.sort_ws = p + off; off += align_up(r.sort_bytes, 256);
  r.total = off;
  return r;
}

inline size_t rulebook_workspace_bytes(int64_t n, int reach) { return carve_rulebook_ws(nullptr, n, reach).total; }

inline int grid_for(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, 256);
  return (int)(b > kNumSMs * 8 ? kNumSMs * 8 : b);
}

inline int check_geometry(const char* who, Geometry& g, const int* shape, const int* k, const int* s, const int* p,
                          const int* d) {
  g.kv = 1;
  for (int a = 0; a < 3; ++a) {
    g.shape[a] = shape[a]; g.k[a] = k[a]; g.s[a] = s ? s[a] : 1; g.d[a] = d ? d[a] : 1; g.p[a] = p ? p[a] : 0;
    B2PC_CHECK_ARG(g.shape[a] > 0 && g.k[a] > 0 && g.s[a] > 0 && g.d[a] > 0 && g.p[a] >= 0, "%s: bad geometry on axis %d", who, a);
    g.oshape[a] = (g.shape[a] + 2 * g.p[a] - g.d[a] * (g.k[a] - 1) - 1) / g.s[a] + 1;
    g.kv *= g.k[a];
  }
  B2PC_CHECK_ARG(g.kv <= 343, "%s: kernel volume %d > 343", who, g.kv);
  return B2PC_OK;
}

inline int launch_rulebook_subm(const int32_t* indices, int64_t n, const int* shape, const int* ksize, const int* dilation,
                                int32_t* pair, void* ws, size_t ws_bytes, cudaStream_t stream) {
  Geometry g;
  int rc = check_geometry("rulebook_subm", g, shape, ksize, nullptr, nullptr, dilation);
  if (rc) return rc;
  B2PC_CHECK_ARG(n >= 0 && n < (1ll << 31), "rulebook_subm: n out of range");
  if (ws_bytes < rulebook_workspace_bytes(n, 1)) { set_error("rulebook_subm: workspace too small"); return B2PC_ERR_WORKSPACE; }
  if (n == 0) return B2PC_OK;
  RulebookWs r = carve_rulebook_ws(ws, n, 1);
  const uint32_t mask = (uint32_t)(r.slots - 1);
  hash_clear_kernel<<<grid_for(r.slots), 256, 0, stream>>>(r.keys, r.slots);
  fill_i32_kernel<<<grid_for(r.slots), 256, 0, stream>>>(r.vals, r.slots, 0x7FFFFFFF);
  hash_insert_rows_kernel<<<grid_for(n), 256, 0, stream>>>((const int4*)indices, n, g, r.keys, r.vals, mask);
  subm_pairs_kernel<<<grid_for(n), 256, 0, stream>>>((const int4*)indices, n, g, r.keys, r.vals, mask, pair);
  count_launches(4);
  B2PC_CHECK_LAUNCH("rulebook_subm");
  return B2PC_OK;
}

inline int launch_rulebook_strided_begin(const int32_t* indices, int64_t n, const int* shape, const int* ksize,
                                         const int* stride, const int* padding, const int* dilation, int64_t* num_out,
                                         void* ws, size_t ws_bytes, cudaStream_t stream) {
  Geometry g;
  int rc = check_geometry("rulebook_strided", g, shape, ksize, stride, padding, dilation);
  if (rc) return rc;
  const int reach = strided_reach(g.k, g.s, g.d);
  B2PC_CHECK_ARG(n >= 0 && n * (int64_t)reach < (1ll << 31), "rulebook_strided: n * reach out of range");
  if (ws_bytes < rulebook_workspace_bytes(n, reach)) { set_error("rulebook_strided: workspace too small"); return B2PC_ERR_WORKSPACE; }
  cudaMemsetAsync(num_out, 0, 2 * sizeof(int64_t), stream);
  if (n == 0) return B2PC_OK;
  RulebookWs r = carve_rulebook_ws(ws, n, reach);
  hash_clear_kernel<<<grid_for(r.slots), 256, 0, stream>>>(r.keys, r.slots);
  strided_collect_kernel<<<grid_for(n), 256, 0, stream>>>((const int4*)indices, n, g, r.keys, (uint32_t)(r.slots - 1), r.uniq,
                                                          (unsigned long long*)num_out);
  count_launches(2);
  B2PC_CHECK_LAUNCH("rulebook_strided_begin");
  return B2PC_OK;
}

inline int launch_rulebook_strided_finish(const int32_t* indices, int64_t n, const int* shape, const int* ksize,
                                          const int* stride, const int* padding, const int* dilation, int64_t m, int batch_count,
                                          int32_t* out_indices, int32_t* pair_fwd, int32_t* pair_bwd, void* ws,
                                          size_t ws_bytes, cudaStream_t stream) {
  Geometry g;
  int rc = check_geometry("rulebook_strided", g, shape, ksize, stride, padding, dilation);
  if (rc) return rc;
  const int reach = strided_reach(g.k, g.s, g.d);
  if (ws_bytes < rulebook_workspace_bytes(n, reach)) { set_error("rulebook_strided: workspace too small"); return B2PC_ERR_WORKSPACE; }
  B2PC_CHECK_ARG(m >= 0 && m <= n * (int64_t)reach, "rulebook_strided_finish: num_out out of range");
  if (n == 0 || m == 0) return B2PC_OK;
  RulebookWs r = carve_rulebook_ws(ws, n, reach);
  const uint32_t mask = (uint32_t)(r.slots - 1);
  // ascending linearised (b,x,y,z): sort the distinct keys on their significant bits
  // key < batch_count * prod(oshape); batch_count comes back with the output count (0 = unknown: allow 2^20 scenes)
  int key_bits = 0;
  {
    const unsigned __int128 nb = batch_count > 0 ? (unsigned __int128)batch_count : ((unsigned __int128)1 << 20);
    unsigned __int128 bound = (unsigned __int128)g.oshape[0] * g.oshape[1] * g.oshape[2] * nb;
    while (key_bits < 64 && (bound >> key_bits) != 0) ++key_bits;
  }
  rc = launch_sort((const int64_t*)r.uniq, m, 1, key_bits, r.order, r.inverse, r.sort_ws, r.sort_bytes, stream);
  if (rc) return rc;
  strided_assign_kernel<<<grid_for(m), 256, 0, stream>>>(r.order, r.uniq, m, g, r.keys, r.vals, mask, out_indices);
  fill_i32_kernel<<<grid_for(m * g.kv), 256, 0, stream>>>(pair_fwd, m * (int64_t)g.kv, -1);
  strided_pairs_kernel<<<grid_for(n), 256, 0, stream>>>((const int4*)indices, n, m, g, r.keys, r.vals, mask, pair_fwd, pair_bwd);
  count_launches(3);
  B2PC_CHECK_LAUNCH("rulebook_strided_finish");
  return B2PC_OK;
}

}  // namespace b2pc
