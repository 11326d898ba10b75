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
