// GPU GridSample (voxelisation) for a whole batch of raw scenes: replaces the per-scene numpy pipeline of
// pointcept/datasets/transform.py:840-958 (floor(coord / grid_size), FNV-1a / ravel hash :983-1011, argsort, np.unique with
// inverse + counts, per-voxel pick) and delivers the result already in collate_fn's layout (datasets/utils.py:19-73): one
// concatenated point list with a cumulative `offset`.
//
// Data flow (all scenes at once, no per-scene launches):
//   gs_extent_kernel   per-scene min / max of floor(coord / grid)                       20 B in per point
//   gs_key_kernel      grid_coord - min, hash -> keys[b][r] (one padded row per scene)   12 B in, 24 + 8 B out
//   launch_sort        stable 8-bit LSD radix sort of every row (sort.cuh)               8 passes of 64-bit keys
//   gs_count / pool_plan_scan / gs_scatter   run heads of the sorted keys -> voxel ids, run starts, idx_sort
//   gs_finish_kernel   inverse (rank of the voxel inside its scene), counts, per-scene count.max, new offsets
//   gs_select_kernel   one point per voxel: train (random member) or test (member f % count, fragment f)
// Integer results (grid_coord, inverse, counts, voxel order = ascending hash inside a scene) are exactly numpy's; which member of
// a voxel comes first is implementation-defined in the reference (np.argsort's default sort is not stable) and is the stable
// order here.
#pragma once
#include "common.cuh"
#include "sort.cuh"
#include "pool.cuh"

namespace b2pc {

constexpr int kGsThreads = 256;

// scene of global point i: first b with offset[b] > i
__device__ __forceinline__ int gs_scene_of(const int64_t* __restrict__ offset, int nb, int64_t i) {
  int lo = 0, hi = nb - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (offset[mid] > i) hi = mid; else lo = mid + 1;
  }
  return lo;
}

struct GsGrid { double g[3]; float gf[3]; };

// floor(coord / grid) the way numpy computes it: in float64 (NumPy >= 2: float32 array / 0-d float64 array promotes) or in
// float32 (NumPy 1.x value-based casting keeps float32 and rounds the grid size to float32 first)
template <bool F64>
__device__ __forceinline__ void gs_cell(const float* __restrict__ c, const GsGrid& gg, int64_t g[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (F64) g[j] = (int64_t)floor((double)c[j] / gg.g[j]);
    else g[j] = (int64_t)floorf(__fdiv_rn(c[j], gg.gf[j]));
  }
}

// ext[b][0..2] = min, ext[b][3..5] = max (int64, pre-set to +inf / -inf by gs_init_kernel)
__global__ void __launch_bounds__(kGsThreads)
gs_init_kernel(int64_t* __restrict__ ext, int nb, int64_t* __restrict__ meta, int n_meta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb * 6) ext[i] = (i % 6) < 3 ? INT64_MAX : INT64_MIN;
  if (i < n_meta) meta[i] = 0;
}

template <bool F64>
__global__ void __launch_bounds__(kGsThreads)
gs_extent_kernel(const float* __restrict__ coord, const int64_t* __restrict__ offset, int nb, int64_t n, GsGrid gg,
                 int64_t* __restrict__ ext) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool ok = i < n;
  int64_t g[3] = {0, 0, 0};
  int b = -1;
  if (ok) {
    b = gs_scene_of(offset, nb, i);
    gs_cell<F64>(coord + i * 3, gg, g);
  }
  // a warp almost always lies inside one scene: reduce there, one atomic per warp and component
  const uint32_t same = __match_any_sync(0xFFFFFFFFu, b);
  if (same == 0xFFFFFFFFu) {
    if (b < 0) return;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int64_t lo = g[j], hi = g[j];
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const int64_t l2 = __shfl_xor_sync(0xFFFFFFFFu, lo, o), h2 = __shfl_xor_sync(0xFFFFFFFFu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
      }
      if ((threadIdx.x & 31) == 0) {
        atomicMin((long long*)&ext[b * 6 + j], (long long)lo);
        atomicMax((long long*)&ext[b * 6 + 3 + j], (long long)hi);
      }
    }
  } else if (ok) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      atomicMin((long long*)&ext[b * 6 + j], (long long)g[j]);
      atomicMax((long long*)&ext[b * 6 + 3 + j], (long long)g[j]);
    }
  }
}

// keys[b * L + r] for point r of scene b (rows pre-filled with ~0 so the padding sorts last and, the sort being stable, after
// every real key); grid_coord[i] = cell - min (int64, numpy's astype(int))
template <bool F64>
__global__ void __launch_bounds__(kGsThreads)
gs_key_kernel(const float* __restrict__ coord, const int64_t* __restrict__ offset, int nb, int64_t n, int64_t row_len, GsGrid gg,
              int hash_type, const int64_t* __restrict__ ext, int64_t* __restrict__ grid_coord, uint64_t* __restrict__ keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = gs_scene_of(offset, nb, i);
  const int64_t start = b ? offset[b - 1] : 0;
  int64_t g[3];
  gs_cell<F64>(coord + i * 3, gg, g);
  uint64_t a[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    g[j] -= ext[b * 6 + j];
    a[j] = (uint64_t)g[j];
    grid_coord[i * 3 + j] = g[j];
  }
  uint64_t key;
  if (hash_type == 0) {   // FNV-1a 64 over the three uint64 lanes, transform.py:997-1011
    key = 14695981039346656037ull;
#pragma unroll
    for (int j = 0; j < 3; ++j) { key *= 1099511628211ull; key ^= a[j]; }
  } else {                // ravel, transform.py:980-995: ((a0 * m1) + a1) * m2 + a2 with m = max + 1 of the zero-based cells
    const uint64_t m1 = (uint64_t)(ext[b * 6 + 4] - ext[b * 6 + 1]) + 1, m2 = (uint64_t)(ext[b * 6 + 5] - ext[b * 6 + 2]) + 1;
    key = (a[0] * m1 + a[1]) * m2 + a[2];
  }
  keys[(int64_t)b * row_len + (i - start)] = key;
}

__device__ __forceinline__ bool gs_is_head(const uint64_t* __restrict__ keys, const int64_t* __restrict__ order, int64_t row_len,
                                           const int64_t* __restrict__ offset, int nb, int64_t p, int* b_out, int64_t* r_out) {
  const int b = gs_scene_of(offset, nb, p);
  const int64_t r = p - (b ? offset[b - 1] : 0);
  *b_out = b;
  *r_out = r;
  if (r == 0) return true;
  const int64_t base = (int64_t)b * row_len;
  return keys[base + order[base + r]] != keys[base + order[base + r - 1]];
}

__global__ void __launch_bounds__(kPpThreads)
gs_count_kernel(const uint64_t* __restrict__ keys, const int64_t* __restrict__ order, int64_t row_len, const int64_t* __restrict__ offset,
                int nb, int64_t n, uint32_t* __restrict__ block_counts) {
  __shared__ uint32_t wsum[kPpThreads / 32];
  uint32_t cnt = 0;
  const int64_t base = (int64_t)blockIdx.x * kPpTile;
#pragma unroll
  for (int it = 0; it < kPpItems; ++it) {
    const int64_t p = base + it * kPpThreads + threadIdx.x;
    int b; int64_t r;
    if (p < n && gs_is_head(keys, order, row_len, offset, nb, p, &b, &r)) ++cnt;
  }
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kPpThreads / 32; ++w) t += wsum[w];
    block_counts[blockIdx.x] = t;
  }
}

// p = position in the concatenation of the sorted scenes.  vox_of_sorted[p] = global voxel id, vox_start[v] = p of its head,
// sort_index[p] = global row of the p-th sorted point (the reference's idx_sort, scene by scene), first_vox[b] = id of scene
// b's first voxel.
__global__ void __launch_bounds__(kPpThreads)
gs_scatter_kernel(const uint64_t* __restrict__ keys, const int64_t* __restrict__ order, int64_t row_len, const int64_t* __restrict__ offset,
                  int nb, int64_t n, const uint32_t* __restrict__ block_offsets, int32_t* __restrict__ vox_of_sorted,
                  int64_t* __restrict__ vox_start, int64_t* __restrict__ sort_index, int64_t* __restrict__ first_vox) {
  __shared__ uint32_t wsum[kPpItems][kPpThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kPpTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool head[kPpItems];
  uint32_t incl[kPpItems];
  int sb[kPpItems];
  int64_t sr[kPpItems];
#pragma unroll
  for (int it = 0; it < kPpItems; ++it) {
    const int64_t p = base + it * kPpThreads + threadIdx.x;
    sb[it] = 0; sr[it] = 0;
    head[it] = p < n && gs_is_head(keys, order, row_len, offset, nb, p, &sb[it], &sr[it]);
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, head[it]);
    incl[it] = __popc(bal & (0xFFFFFFFFu >> (31 - lane)));
    if (lane == 31) wsum[it][warp] = incl[it];
  }
  __syncthreads();
  uint32_t run = block_offsets[blockIdx.x];
#pragma unroll
  for (int it = 0; it < kPpItems; ++it) {
    uint32_t before = run;
    for (int w = 0; w < kPpThreads / 32; ++w) {
      if (w < warp) before += wsum[it][w];
      run += wsum[it][w];
    }
    const int64_t p = base + it * kPpThreads + threadIdx.x;
    if (p < n) {
      const int64_t v = (int64_t)(before + incl[it]) - 1;
      const int64_t start = p - sr[it];
      vox_of_sorted[p] = (int32_t)v;
      sort_index[p] = start + order[(int64_t)sb[it] * row_len + sr[it]];
      if (head[it]) {
        vox_start[v] = p;
        if (sr[it] == 0) first_vox[sb[it]] = v;
      }
    }
  }
}

// meta layout (int64): [0] = M (voxels in the batch), [1 .. B] = cumulative voxel offsets (the collated `offset` after
// sampling), [1+B .. 2B] = count.max() of every scene, [1+2B .. 1+2B+3B) = per-scene minimum cell (min_coord / grid_size)
__global__ void __launch_bounds__(kGsThreads)
gs_finish_kernel(const int32_t* __restrict__ vox_of_sorted, const int64_t* __restrict__ vox_start, const int64_t* __restrict__ sort_index,
                 const int64_t* __restrict__ first_vox, const int64_t* __restrict__ offset, int nb, int64_t n, const int64_t* __restrict__ total,
                 const int64_t* __restrict__ ext, int64_t* __restrict__ inverse, int64_t* __restrict__ vox_count, int64_t* __restrict__ meta) {
  const int64_t m = *total;
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p == 0) meta[0] = m;
  if (p < nb) {
    meta[1 + p] = (p + 1 < nb ? first_vox[p + 1] : m);
#pragma unroll
    for (int j = 0; j < 3; ++j) meta[1 + 2 * nb + 3 * p + j] = ext[p * 6 + j];
  }
  if (p >= n) return;
  const int b = gs_scene_of(offset, nb, p);
  const int64_t v = vox_of_sorted[p];
  inverse[sort_index[p]] = v - first_vox[b];
  if (vox_start[v] == p) {
    const int64_t cnt = (v + 1 < m ? vox_start[v + 1] : n) - p;
    vox_count[v] = cnt;
    atomicMax((long long*)&meta[1 + nb + b], (long long)cnt);
  }
}

__device__ __forceinline__ uint64_t gs_mix(uint64_t x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// mode 0 (train, transform.py:876-881): member (u % count) with u uniform in [0, count.max()) -- the reference's
// `np.random.randint(0, count.max(), count.size) % count`, modulo bias included; mode 1 (test, :914-916): member (arg % count).
__global__ void __launch_bounds__(kGsThreads)
gs_select_kernel(const int64_t* __restrict__ sort_index, const int64_t* __restrict__ vox_start, const int64_t* __restrict__ vox_count,
                 const int64_t* __restrict__ meta, int nb, int64_t m, int mode, uint64_t arg, int64_t* __restrict__ idx_out) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= m) return;
  const int64_t cnt = vox_count[v];
  int64_t pick;
  if (mode == 0) {
    const int b = gs_scene_of(meta + 1, nb, v);
    const uint64_t cmax = (uint64_t)meta[1 + nb + b];
    const uint64_t u = gs_mix(arg ^ gs_mix((uint64_t)v)) % cmax;
    pick = (int64_t)(u % (uint64_t)cnt);
  } else {
    pick = (int64_t)(arg % (uint64_t)cnt);
  }
  idx_out[v] = sort_index[vox_start[v] + pick];
}

// dst[i, :] = src[idx[i], :] for rows of row_bytes bytes (index_operator of the transforms / collate; any payload dtype)
template <typename V>
__global__ void __launch_bounds__(kGsThreads)
gather_rows_kernel(const V* __restrict__ src, int64_t row_vecs, const int64_t* __restrict__ idx, int64_t m, V* __restrict__ dst) {
  const int64_t total = m * row_vecs;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / row_vecs, c = e - i * row_vecs;
    dst[e] = src[idx[i] * row_vecs + c];
  }
}

// displacement of the picked point to its cell centre, transform.py:893-903: (coord / grid - min) - grid_coord - 0.5
template <bool F64, typename O>
__global__ void __launch_bounds__(kGsThreads)
gs_displacement_kernel(const float* __restrict__ coord, const int64_t* __restrict__ idx, const int64_t* __restrict__ meta, int nb, int64_t m,
                       GsGrid gg, O* __restrict__ out) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= m) return;
  const int b = gs_scene_of(meta + 1, nb, v);
  const int64_t* mn = meta + 1 + 2 * nb + 3 * b;
  const float* c = coord + idx[v] * 3;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (F64) {
      const double s = (double)c[j] / gg.g[j];
      const double cell = floor(s) - (double)mn[j];
      out[v * 3 + j] = (O)(((s - (double)mn[j]) - cell) - 0.5);
    } else {
      // NumPy 1.x: `scaled_coord -= min_coord` rounds the difference to float32 in place, the final expression promotes to float64
      const float s = __fdiv_rn(c[j], gg.gf[j]);
      const float sm = s - (float)mn[j];
      const double cell = (double)((int64_t)floorf(s) - mn[j]);
      out[v * 3 + j] = (O)(((double)sm - cell) - 0.5);
    }
  }
}

struct GsWorkspace {
  uint64_t* keys; int64_t* order; int64_t* inv; int64_t* ext; int64_t* first_vox; int32_t* vox_of_sorted; uint32_t* counts; void* sort_ws;
  size_t sort_ws_bytes, total;
};

inline GsWorkspace gs_workspace(void* ws, int64_t n, int nb, int64_t row_len) {
  GsWorkspace w;
  char* p = (char*)ws;
  const size_t rows = (size_t)nb * (size_t)row_len;
  auto take = [&](size_t bytes) { char* q = p; p += align_up(bytes, 256); return q; };
  w.keys = (uint64_t*)take(rows * 8);
  w.order = (int64_t*)take(rows * 8);
  w.inv = (int64_t*)take(rows * 8);
  w.ext = (int64_t*)take((size_t)nb * 6 * 8);
  w.first_vox = (int64_t*)take((size_t)nb * 8);
  w.vox_of_sorted = (int32_t*)take((size_t)(n > 0 ? n : 1) * 4);
  w.counts = (uint32_t*)take(((size_t)ceil_div(n > 0 ? n : 1, kPpTile) + 1) * 4);
  w.sort_ws_bytes = sort_workspace_bytes(row_len, nb);
  w.sort_ws = take(w.sort_ws_bytes);
  w.total = (size_t)(p - (char*)ws);
  return w;
}

inline size_t grid_sample_workspace_bytes(int64_t n, int nb, int64_t row_len) { return gs_workspace(nullptr, n, nb, row_len).total; }

inline int launch_grid_sample_plan(const float* coord, const int64_t* offset, int nb, int64_t n, int64_t row_len, const double* grid_host,
                                   int hash_type, int math_f64, int64_t* grid_coord, int64_t* inverse, int64_t* sort_index,
                                   int64_t* vox_start, int64_t* vox_count, int64_t* meta, void* ws, size_t ws_bytes, cudaStream_t stream) {
  B2PC_CHECK_ARG(n > 0 && n < (1ll << 31) && nb >= 1 && nb <= 4096 && row_len >= 1 && row_len <= n, "grid_sample: bad sizes");
  B2PC_CHECK_ARG(hash_type == 0 || hash_type == 1, "grid_sample: hash_type must be 0 (fnv) or 1 (ravel)");
  B2PC_CHECK_ARG(grid_host[0] > 0 && grid_host[1] > 0 && grid_host[2] > 0, "grid_sample: grid_size must be positive");
  if (ws_bytes < grid_sample_workspace_bytes(n, nb, row_len)) { set_error("grid_sample: workspace too small"); return B2PC_ERR_WORKSPACE; }
  const GsWorkspace w = gs_workspace(ws, n, nb, row_len);
  GsGrid gg;
  for (int j = 0; j < 3; ++j) { gg.g[j] = grid_host[j]; gg.gf[j] = (float)grid_host[j]; }
  const int n_meta = 1 + 5 * nb;
  const int nblk = (int)ceil_div(n, kGsThreads);
  gs_init_kernel<<<(int)ceil_div(nb * 6 > n_meta ? nb * 6 : n_meta, kGsThreads), kGsThreads, 0, stream>>>(w.ext, nb, meta, n_meta);
  cudaMemsetAsync(w.keys, 0xFF, (size_t)nb * row_len * 8, stream);
  if (math_f64) {
    gs_extent_kernel<true><<<nblk, kGsThreads, 0, stream>>>(coord, offset, nb, n, gg, w.ext);
    gs_key_kernel<true><<<nblk, kGsThreads, 0, stream>>>(coord, offset, nb, n, row_len, gg, hash_type, w.ext, grid_coord, w.keys);
  } else {
    gs_extent_kernel<false><<<nblk, kGsThreads, 0, stream>>>(coord, offset, nb, n, gg, w.ext);
    gs_key_kernel<false><<<nblk, kGsThreads, 0, stream>>>(coord, offset, nb, n, row_len, gg, hash_type, w.ext, grid_coord, w.keys);
  }
  count_launches(3);
  const int rc = launch_sort((const int64_t*)w.keys, row_len, nb, 64, w.order, w.inv, w.sort_ws, w.sort_ws_bytes, stream);
  if (rc != B2PC_OK) return rc;
  const int pblk = (int)ceil_div(n, kPpTile);
  gs_count_kernel<<<pblk, kPpThreads, 0, stream>>>(w.keys, w.order, row_len, offset, nb, n, w.counts);
  // exclusive scan of the block counts; the total lands in first_vox-independent scratch: counts[pblk]
  pool_plan_scan_kernel<<<1, 1024, 0, stream>>>(w.counts, pblk, w.inv, 0);   // total -> w.inv[0] (int64 scratch, free after the sort)
  gs_scatter_kernel<<<pblk, kPpThreads, 0, stream>>>(w.keys, w.order, row_len, offset, nb, n, w.counts, w.vox_of_sorted, vox_start, sort_index,
                                                     w.first_vox);
  gs_finish_kernel<<<nblk, kGsThreads, 0, stream>>>(w.vox_of_sorted, vox_start, sort_index, w.first_vox, offset, nb, n,
                                                    w.inv, w.ext, inverse, vox_count, meta);
  count_launches(4);
  B2PC_CHECK_LAUNCH("grid_sample_plan");
  return B2PC_OK;
}

}  // namespace b2pc
