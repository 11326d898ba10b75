// Serialized pooling (reference point_transformer_v3m1_base.py:371-444): clusters are runs of the
// order-0 sorted sequence, so the [indices] gather, torch_scatter.segment_csr(max) and its backward collapse into one
// kernel each.  HBM-bound: forward reads N*C + writes M*C (+4 B argmax per output), backward writes N*C once.
#pragma once
#include "common.cuh"

namespace b2pc {

// out[m, c] = max_{r in [start[m], start[m]+len[m])} x[order[r], c];  arg[m, c] = winning source row (first maximum)
template <typename T>
__global__ void __launch_bounds__(256)
segment_max_fwd_kernel(const T* __restrict__ x, const int64_t* __restrict__ order, const int64_t* __restrict__ start,
                       const int64_t* __restrict__ len, int64_t m, int c, T* __restrict__ out, int32_t* __restrict__ arg) {
  const int64_t total = m * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t seg = i / c;
    const int ch = (int)(i % c);
    const int64_t s = start[seg], e = s + len[seg];
    float best = -INFINITY;
    int64_t best_row = order[s];
    for (int64_t r = s; r < e; ++r) {
      const int64_t row = order[r];
      const float v = to_f32(x[row * c + ch]);
      if (v > best) { best = v; best_row = row; }
    }
    out[i] = from_f32<T>(best);
    arg[i] = (int32_t)best_row;
  }
}

// dx[arg[m, c], c] = dout[m, c]   (dx pre-zeroed; targets are unique per channel)
template <typename T>
__global__ void __launch_bounds__(256)
segment_max_bwd_kernel(const T* __restrict__ dout, const int32_t* __restrict__ arg, int64_t m, int c, T* __restrict__ dx) {
  const int64_t total = m * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    dx[(int64_t)arg[i] * c + (i % c)] = dout[i];
}

template <typename T>
inline int launch_segment_max_fwd_t(const void* x, const int64_t* order, const int64_t* start, const int64_t* len, int64_t m, int c,
                                    void* out, int32_t* arg, cudaStream_t stream) {
  int64_t b = ceil_div(m * c, 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  segment_max_fwd_kernel<T><<<(int)b, 256, 0, stream>>>((const T*)x, order, start, len, m, c, (T*)out, arg);
  return 0;
}
template <typename T>
inline int launch_segment_max_bwd_t(const void* dout, const int32_t* arg, int64_t m, int c, void* dx, cudaStream_t stream) {
  int64_t b = ceil_div(m * c, 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  segment_max_bwd_kernel<T><<<(int)b, 256, 0, stream>>>((const T*)dout, arg, m, c, (T*)dx);
  return 0;
}

inline int launch_segment_max_fwd(const void* x, int dtype, const int64_t* order, const int64_t* start, const int64_t* len, int64_t m,
                                  int c, void* out, int32_t* arg, cudaStream_t stream) {
  if (m == 0) return B2PC_OK;
  switch (dtype) {
    case B2PC_F32: launch_segment_max_fwd_t<float>(x, order, start, len, m, c, out, arg, stream); break;
    case B2PC_F16: launch_segment_max_fwd_t<__half>(x, order, start, len, m, c, out, arg, stream); break;
    case B2PC_BF16: launch_segment_max_fwd_t<__nv_bfloat16>(x, order, start, len, m, c, out, arg, stream); break;
    default: set_error("segment_max_fwd: unknown dtype %d", dtype); return B2PC_ERR_INVALID_ARG;
  }
  count_launches(1);
  B2PC_CHECK_LAUNCH("segment_max_fwd");
  return B2PC_OK;
}

inline int launch_segment_max_bwd(const void* dout, int dtype, const int32_t* arg, int64_t m, int c, int64_t n, void* dx,
                                  cudaStream_t stream) {
  const size_t es = dtype == B2PC_F32 ? 4 : 2;
  cudaMemsetAsync(dx, 0, (size_t)n * c * es, stream);
  if (m == 0) return B2PC_OK;
  switch (dtype) {
    case B2PC_F32: launch_segment_max_bwd_t<float>(dout, arg, m, c, dx, stream); break;
    case B2PC_F16: launch_segment_max_bwd_t<__half>(dout, arg, m, c, dx, stream); break;
    case B2PC_BF16: launch_segment_max_bwd_t<__nv_bfloat16>(dout, arg, m, c, dx, stream); break;
    default: set_error("segment_max_bwd: unknown dtype %d", dtype); return B2PC_ERR_INVALID_ARG;
  }
  count_launches(1);
  B2PC_CHECK_LAUNCH("segment_max_bwd");
  return B2PC_OK;
}

}  // namespace b2pc

namespace b2pc {
// ---------------------------------------------------------------------------------------------------------------------
// Index side of SerializedPooling (point_transformer_v3m1_base.py:371-398: code >> 3*depth, torch.unique, argsort of the
// cluster ids, head indices, pooled codes / grid coordinates / batch ids) as three small kernels and ONE deferred host read:
// clusters are the runs of equal (code[0] >> 3*pd) in the order-0 sorted sequence.
//   pool_plan_count_kernel    flags run heads, per-block head counts
//   pool_plan_scan_kernel     exclusive scan of the block counts (one block), total M -> meta[0]
//   pool_plan_scatter_kernel  cluster[row], head_pos / head_indices, pooled codes of every order, pooled grid / batch,
//                             per-scene counts -> meta[1 + b]
//   pool_plan_lengths_kernel  lengths[c] = head_pos[c + 1] - head_pos[c]
// All outputs are sized for the upper bound N; the caller slices them to M after reading meta (B + 1 int64).
constexpr int kPpThreads = 256;
constexpr int kPpItems = 4;
constexpr int kPpTile = kPpThreads * kPpItems;

__device__ __forceinline__ bool pp_is_head(const int64_t* __restrict__ code0, const int64_t* __restrict__ order0, int64_t i, int sh) {
  if (i == 0) return true;
  return (code0[order0[i]] >> sh) != (code0[order0[i - 1]] >> sh);
}

__global__ void __launch_bounds__(kPpThreads)
pool_plan_count_kernel(const int64_t* __restrict__ code0, const int64_t* __restrict__ order0, int64_t n, int sh,
                       uint32_t* __restrict__ block_counts) {
  __shared__ uint32_t wsum[kPpThreads / 32];
  uint32_t cnt = 0;
  const int64_t base = (int64_t)blockIdx.x * kPpTile;
#pragma unroll
  for (int it = 0; it < kPpItems; ++it) {
    const int64_t i = base + it * kPpThreads + threadIdx.x;
    if (i < n && pp_is_head(code0, order0, i, sh)) ++cnt;
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

// exclusive scan in place, total -> meta[0]; meta[1..] (scene counts) zeroed
__global__ void __launch_bounds__(1024)
pool_plan_scan_kernel(uint32_t* __restrict__ block_counts, int nblocks, int64_t* __restrict__ meta, int n_scene) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  for (int b = threadIdx.x; b < n_scene; b += 1024) meta[1 + b] = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_counts[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u;
    const uint32_t carry = carry_s;
    if (i < nblocks) block_counts[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) meta[0] = carry_s;
}

// element order inside a tile is (item, thread): element e = it * kPpThreads + tid, i.e. ascending i over it then tid
__global__ void __launch_bounds__(kPpThreads)
pool_plan_scatter_kernel(const int64_t* __restrict__ code, int n_orders, int64_t n, const int64_t* __restrict__ order0,
                         const int64_t* __restrict__ batch, const int32_t* __restrict__ grid, int pd,
                         const uint32_t* __restrict__ block_offsets, int64_t* __restrict__ cluster, int64_t* __restrict__ head_pos,
                         int64_t* __restrict__ head_indices, int64_t* __restrict__ code_out, int64_t cap, int64_t* __restrict__ batch_out,
                         int32_t* __restrict__ grid_out, int64_t* __restrict__ meta) {
  __shared__ uint32_t wsum[kPpItems][kPpThreads / 32];
  const int sh = 3 * pd;
  const int64_t base = (int64_t)blockIdx.x * kPpTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool head[kPpItems];
  uint32_t incl[kPpItems];
#pragma unroll
  for (int it = 0; it < kPpItems; ++it) {
    const int64_t i = base + it * kPpThreads + threadIdx.x;
    head[it] = i < n && pp_is_head(code, order0, i, sh);
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, head[it]);
    incl[it] = __popc(b & (0xFFFFFFFFu >> (31 - lane)));        // heads at lanes <= this one
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
    const int64_t i = base + it * kPpThreads + threadIdx.x;
    if (i < n) {
      const int64_t row = order0[i];
      const int64_t cid = (int64_t)(before + incl[it]) - 1;     // cluster of this element
      cluster[row] = cid;
      if (head[it]) {
        head_pos[cid] = i;
        head_indices[cid] = row;
        for (int o = 0; o < n_orders; ++o) code_out[(int64_t)o * cap + cid] = code[(int64_t)o * n + row] >> sh;
        const int64_t b = batch[row];
        batch_out[cid] = b;
        grid_out[cid * 3 + 0] = grid[row * 3 + 0] >> pd;
        grid_out[cid * 3 + 1] = grid[row * 3 + 1] >> pd;
        grid_out[cid * 3 + 2] = grid[row * 3 + 2] >> pd;
        atomicAdd((unsigned long long*)&meta[1 + b], 1ull);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
pool_plan_lengths_kernel(const int64_t* __restrict__ head_pos, const int64_t* __restrict__ meta, int64_t n, int64_t* __restrict__ lengths) {
  const int64_t m = meta[0];
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < m; c += (int64_t)gridDim.x * blockDim.x)
    lengths[c] = (c + 1 < m ? head_pos[c + 1] : n) - head_pos[c];
}

inline size_t pool_plan_workspace_bytes(int64_t n) { return (size_t)ceil_div(n > 0 ? n : 1, kPpTile) * sizeof(uint32_t) + 256; }

inline int launch_pool_plan(const int64_t* code, int n_orders, int64_t n, const int64_t* order0, const int64_t* batch, const int32_t* grid,
                            int pooling_depth, int n_scene, int64_t* cluster, int64_t* head_pos, int64_t* head_indices, int64_t* lengths,
                            int64_t* code_out, int64_t* batch_out, int32_t* grid_out, int64_t* meta, void* ws, size_t ws_bytes,
                            cudaStream_t stream) {
  B2PC_CHECK_ARG(n > 0 && n < (1ll << 31) && n_orders >= 1 && pooling_depth >= 0 && pooling_depth <= 16 && n_scene >= 1, "pool_plan: bad sizes");
  if (ws_bytes < pool_plan_workspace_bytes(n)) { set_error("pool_plan: workspace too small"); return B2PC_ERR_WORKSPACE; }
  const int nblocks = (int)ceil_div(n, kPpTile);
  uint32_t* counts = (uint32_t*)ws;
  pool_plan_count_kernel<<<nblocks, kPpThreads, 0, stream>>>(code, order0, n, 3 * pooling_depth, counts);
  pool_plan_scan_kernel<<<1, 1024, 0, stream>>>(counts, nblocks, meta, n_scene);
  pool_plan_scatter_kernel<<<nblocks, kPpThreads, 0, stream>>>(code, n_orders, n, order0, batch, grid, pooling_depth, counts, cluster, head_pos,
                                                               head_indices, code_out, n, batch_out, grid_out, meta);
  int lb = (int)ceil_div(n, 256);
  if (lb > kNumSMs * 4) lb = kNumSMs * 4;
  pool_plan_lengths_kernel<<<lb, 256, 0, stream>>>(head_pos, meta, n, lengths);
  count_launches(4);
  B2PC_CHECK_LAUNCH("pool_plan");
  return B2PC_OK;
}
}  // namespace b2pc
