// Fused LayerNorm forward / backward over point features [N, C] (SURVEY.md 8(f).2 "memory-bound glue"):
// one pass over x per direction, fp32 statistics, any of fp32 / fp16 / bf16 on either side.  HBM-bound:
// forward reads N*C*sizeof(X) and writes N*C*sizeof(Y) (+8 B/row of statistics); backward reads dy and x once,
// writes dx once, and reduces dgamma/dbeta through per-block partials (deterministic).
// A warp owns a row at a time; lane l holds channels l*V .. l*V+V-1 of every 32*V-channel slab (V = 4), so that C = 32
// still keeps all lanes busy with 1 element... (small C: several rows per warp, see kRowsPerWarp).
#pragma once
#include "common.cuh"

namespace b2pc {

constexpr int kLnThreads = 256;
constexpr int kLnMaxPerLane = 16;   // C <= 512

template <typename T> __device__ __forceinline__ float ln_load(const T* p, int64_t i) { return to_f32(p[i]); }

// rows are distributed over warps; each lane handles channels lane, lane+32, ... (coalesced across the warp)
template <typename X, typename Y>
__global__ void __launch_bounds__(kLnThreads)
layer_norm_fwd_kernel(const X* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t n, int c,
                      float eps, Y* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * (int64_t)kLnThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kLnThreads) >> 5;
  const int per = c >> 5;  // channels per lane (c is a multiple of 32)
  float g[kLnMaxPerLane], b[kLnMaxPerLane];
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i)
    if (i < per) { g[i] = gamma[lane + 32 * i]; b[i] = beta ? beta[lane + 32 * i] : 0.f; }
#pragma unroll 2
  for (int64_t r = warp_global; r < n; r += n_warps) {
    float v[kLnMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
      if (i < per) { v[i] = to_f32(x[r * c + lane + 32 * i]); s += v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    const float mu = s / c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
      if (i < per) { const float d = v[i] - mu; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
    const float rs = rsqrtf(q / c + eps);
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
      if (i < per) y[r * c + lane + 32 * i] = from_f32<Y>((v[i] - mu) * rs * g[i] + b[i]);
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

template <typename X, typename Y>
__global__ void __launch_bounds__(kLnThreads)
layer_norm_bwd_kernel(const Y* __restrict__ dy, const X* __restrict__ x, const float* __restrict__ gamma,
                      const float* __restrict__ mean, const float* __restrict__ rstd, int64_t n, int c, X* __restrict__ dx,
                      float* __restrict__ part_g, float* __restrict__ part_b) {
  __shared__ float red[kLnThreads / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (blockIdx.x * (int64_t)kLnThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kLnThreads) >> 5;
  const int per = c >> 5;
  float g[kLnMaxPerLane], ag[kLnMaxPerLane], ab[kLnMaxPerLane];
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) { ag[i] = 0.f; ab[i] = 0.f; if (i < per) g[i] = gamma[lane + 32 * i]; }
#pragma unroll 2
  for (int64_t r = warp_global; r < n; r += n_warps) {
    const float mu = mean[r], rs = rstd[r];
    float xh[kLnMaxPerLane], dh[kLnMaxPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
      if (i < per) {
        const float d = to_f32(dy[r * c + lane + 32 * i]);
        xh[i] = (to_f32(x[r * c + lane + 32 * i]) - mu) * rs;
        dh[i] = d * g[i];
        s1 += dh[i];
        s2 += dh[i] * xh[i];
        ag[i] += d * xh[i];
        ab[i] += d;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
    const float m1 = s1 / c, m2 = s2 / c;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i)
      if (i < per) dx[r * c + lane + 32 * i] = from_f32<X>(rs * (dh[i] - m1 - xh[i] * m2));
  }
  // block partials: sum over the block's warps, fixed order
  for (int i = 0; i < per; ++i) {
    red[warp][lane] = ag[i];
    __syncthreads();
    if (warp == 0) {
      float t = 0.f;
      for (int w = 0; w < kLnThreads / 32; ++w) t += red[w][lane];
      part_g[(int64_t)blockIdx.x * c + lane + 32 * i] = t;
    }
    __syncthreads();
    red[warp][lane] = ab[i];
    __syncthreads();
    if (warp == 0) {
      float t = 0.f;
      for (int w = 0; w < kLnThreads / 32; ++w) t += red[w][lane];
      part_b[(int64_t)blockIdx.x * c + lane + 32 * i] = t;
    }
    __syncthreads();
  }
}

// one block per 32 channels; 8 warps stride over the per-block partials, then a fixed-order shared-memory sum
__global__ void __launch_bounds__(256)
layer_norm_param_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b, int blocks, int c,
                               float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float sg[8][32], sb[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  float ag = 0.f, ab = 0.f;
  for (int b = grp; b < blocks; b += 8) { ag += part_g[(int64_t)b * c + ch]; ab += part_b[(int64_t)b * c + ch]; }
  sg[grp][lane] = ag; sb[grp][lane] = ab;
  __syncthreads();
  if (grp == 0) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tg += sg[w][lane]; tb += sb[w][lane]; }
    dgamma[ch] = tg;
    if (dbeta) dbeta[ch] = tb;
  }
}

inline int ln_blocks(int64_t n) {
  int64_t b = ceil_div(n, kLnThreads / 32 * 4);
  if (b > kNumSMs * 4) b = kNumSMs * 4;
  return (int)(b < 1 ? 1 : b);
}
inline size_t layer_norm_bwd_workspace_bytes(int64_t n, int c) { return (size_t)ln_blocks(n) * c * 2 * sizeof(float) + 256; }

#define B2PC_LN_DISPATCH(XD, YD, CALL)                                                             \
  if (XD == B2PC_F32 && YD == B2PC_F32) { using X = float; using Y = float; CALL; }                \
  else if (XD == B2PC_BF16 && YD == B2PC_F32) { using X = __nv_bfloat16; using Y = float; CALL; }  \
  else if (XD == B2PC_F16 && YD == B2PC_F32) { using X = __half; using Y = float; CALL; }          \
  else if (XD == B2PC_BF16 && YD == B2PC_BF16) { using X = __nv_bfloat16; using Y = __nv_bfloat16; CALL; } \
  else if (XD == B2PC_F16 && YD == B2PC_F16) { using X = __half; using Y = __half; CALL; }         \
  else if (XD == B2PC_F32 && YD == B2PC_BF16) { using X = float; using Y = __nv_bfloat16; CALL; }  \
  else if (XD == B2PC_F32 && YD == B2PC_F16) { using X = float; using Y = __half; CALL; }          \
  else { set_error("layer_norm: unsupported dtype pair (%d, %d)", XD, YD); return B2PC_ERR_UNSUPPORTED; }

inline int launch_layer_norm_fwd(const void* x, int xd, const float* gamma, const float* beta, int64_t n, int c, float eps, void* y,
                                 int yd, float* mean, float* rstd, cudaStream_t stream) {
  B2PC_CHECK_ARG(c % 32 == 0 && c >= 32 && c <= 32 * kLnMaxPerLane, "layer_norm: channels %d not a multiple of 32 in [32,512]", c);
  if (n == 0) return B2PC_OK;
  const int blocks = ln_blocks(n);
  B2PC_LN_DISPATCH(xd, yd, (layer_norm_fwd_kernel<X, Y><<<blocks, kLnThreads, 0, stream>>>((const X*)x, gamma, beta, n, c, eps, (Y*)y, mean, rstd)));
  count_launches(1);
  B2PC_CHECK_LAUNCH("layer_norm_fwd");
  return B2PC_OK;
}

inline int launch_layer_norm_bwd(const void* dy, int yd, const void* x, int xd, const float* gamma, const float* mean,
                                 const float* rstd, int64_t n, int c, void* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                                 cudaStream_t stream) {
  B2PC_CHECK_ARG(c % 32 == 0 && c >= 32 && c <= 32 * kLnMaxPerLane, "layer_norm: channels %d not a multiple of 32 in [32,512]", c);
  if (ws_bytes < layer_norm_bwd_workspace_bytes(n, c)) { set_error("layer_norm_bwd: workspace too small"); return B2PC_ERR_WORKSPACE; }
  const int blocks = ln_blocks(n);
  float* pg = (float*)ws;
  float* pb = pg + (int64_t)blocks * c;
  if (n == 0) {
    cudaMemsetAsync(dgamma, 0, c * sizeof(float), stream);
    if (dbeta) cudaMemsetAsync(dbeta, 0, c * sizeof(float), stream);
    return B2PC_OK;
  }
  B2PC_LN_DISPATCH(xd, yd, (layer_norm_bwd_kernel<X, Y><<<blocks, kLnThreads, 0, stream>>>((const Y*)dy, (const X*)x, gamma, mean, rstd, n, c, (X*)dx, pg, pb)));
  layer_norm_param_reduce_kernel<<<c / 32, 256, 0, stream>>>(pg, pb, blocks, c, dgamma, dbeta);
  count_launches(2);
  B2PC_CHECK_LAUNCH("layer_norm_bwd");
  return B2PC_OK;
}

}  // namespace b2pc
