// Fused cross-entropy of the segmentation head (pointcept/models/losses/misc.py:13-40: nn.CrossEntropyLoss(reduction="mean",
// ignore_index=-1) on seg_logits [N, n_classes]; called from DefaultSegmentorV2.forward, pointcept/models/default.py:83-90).
// Forward: one pass over the logits -> per-row log-sum-exp, mean loss over the non-ignored rows (two-stage, fixed-order
// reduction).  Backward: dlogits = (softmax - onehot) * dloss / count, recomputed from the logits and the saved log-sum-exp.
// Replaces log_softmax + nll_loss (whose forward reduction is a single-block kernel: 0.21 ms at N = 240 k) and their backwards.
#pragma once
#include "common.cuh"

namespace b2pc {

constexpr int kCeThreads = 256;
constexpr int kCeRowsPerBlock = kCeThreads / 32;

template <typename T>
__global__ void __launch_bounds__(kCeThreads)
cross_entropy_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ target, int64_t n, int c, int64_t ignore_index,
                         float* __restrict__ lse, float* __restrict__ partial /* [gridDim.x][2] */) {
  __shared__ float s_loss[kCeRowsPerBlock], s_cnt[kCeRowsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float loss = 0.f, cnt = 0.f;
  for (int64_t row = blockIdx.x * (int64_t)kCeRowsPerBlock + warp; row < n; row += (int64_t)gridDim.x * kCeRowsPerBlock) {
    const T* x = logits + row * c;
    float mx = -INFINITY;
    for (int j = lane; j < c; j += 32) mx = fmaxf(mx, to_f32(x[j]));
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < c; j += 32) sum += expf(to_f32(x[j]) - mx);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float l = mx + logf(sum);
    if (lane == 0) {
      lse[row] = l;
      const int64_t t = target[row];
      if (t != ignore_index && t >= 0 && t < c) { loss += l - to_f32(x[t]); cnt += 1.f; }
    }
  }
  if (lane == 0) { s_loss[warp] = loss; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < kCeRowsPerBlock; ++w) { a += s_loss[w]; b += s_cnt[w]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
}

// out[0] = mean loss, out[1] = number of contributing rows; partials summed in index order by one warp (fixed order)
__global__ void __launch_bounds__(32)
cross_entropy_finish_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ out) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) { a += partial[2 * i]; b += partial[2 * i + 1]; }
#pragma unroll
  for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
  if (threadIdx.x == 0) { out[0] = (float)(a / b); out[1] = (float)b; }   // no valid row: 0/0 = NaN, as torch returns
}

template <typename T>
__global__ void __launch_bounds__(kCeThreads)
cross_entropy_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ target, const float* __restrict__ lse,
                         const float* __restrict__ grad_loss, const float* __restrict__ loss_count, int64_t n, int c,
                         int64_t ignore_index, T* __restrict__ dlogits) {
  const float g = grad_loss[0] / loss_count[1];
  const int64_t total = n * c;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / c;
    const int j = (int)(e - row * c);
    const int64_t t = target[row];
    float d = 0.f;
    if (t != ignore_index && t >= 0 && t < c) d = (expf(to_f32(logits[e]) - lse[row]) - (j == t ? 1.f : 0.f)) * g;
    dlogits[e] = from_f32<T>(d);
  }
}

inline int ce_blocks(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, kCeRowsPerBlock);
  return (int)(b > kNumSMs * 8 ? kNumSMs * 8 : b);
}
inline size_t cross_entropy_workspace_bytes(int64_t n) { return (size_t)ce_blocks(n) * 2 * sizeof(float) + 256; }

}  // namespace b2pc
