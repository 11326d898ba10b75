// Fused LayerNorm forward / backward over point features [N, C] (SURVEY.md 8(f).2 "memory-bound glue"):
// one pass over x per direction, fp32 statistics, any of fp32 / fp16 / bf16 on either side.  HBM-bound:
// forward reads N*C*sizeof(X) and writes N*C*sizeof(Y) (+8 B/row of statistics); backward reads dy and x once,
// writes dx once, and reduces dgamma/dbeta through per-block partials (deterministic).
// Lanes own float4 chunks of a row; narrow rows share a warp (C = 32: 4 rows per warp) so every access is coalesced.
#pragma once
#include "common.cuh"

namespace b2pc {

constexpr int kLnThreads = 256;

// ---- vector access: 4 consecutive channels per lane -----------------------------------------------------------------
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<__half>(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// Mapping: L = min(32, C/4) lanes share a row (each lane owns float4 chunks lane, lane+L, ...: V = C/(4L) of them), so a
// warp covers 32/L rows with fully coalesced 16-byte (fp32) / 8-byte (16-bit) accesses, and the row reductions are
// log2(L) shuffles.
template <typename X, typename Y, int V>
__global__ void __launch_bounds__(kLnThreads)
layer_norm_fwd_kernel(const X* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t n, int c,
                      float eps, Y* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  const int L = c / (4 * V);             // lanes per row (power of two <= 32)
  const int rpw = 32 / L;                // rows per warp
  const int lane = threadIdx.x & 31, sub = lane % L, rin = lane / L;
  const int64_t warp_global = (blockIdx.x * (int64_t)kLnThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kLnThreads) >> 5;
  float4 g[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    g[i] = ld4<float>(gamma + 4 * (sub + i * L));
    b[i] = beta ? ld4<float>(beta + 4 * (sub + i * L)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.f / c;
  for (int64_t r0 = warp_global * rpw; r0 < n; r0 += n_warps * rpw) {
    const int64_t r = r0 + rin;
    const bool ok = r < n;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = ok ? ld4<X>(x + r * c + 4 * (sub + i * L)) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    const float mu = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    for (int o = L >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
    const float rs = rsqrtf(q * inv_c + eps);
    if (ok) {
#pragma unroll
      for (int i = 0; i < V; ++i)
        st4<Y>(y + r * c + 4 * (sub + i * L),
               make_float4((v[i].x - mu) * rs * g[i].x + b[i].x, (v[i].y - mu) * rs * g[i].y + b[i].y,
                           (v[i].z - mu) * rs * g[i].z + b[i].z, (v[i].w - mu) * rs * g[i].w + b[i].w));
      if (sub == 0) { mean[r] = mu; rstd[r] = rs; }
    }
  }
}

template <typename X, typename Y, int V>
__global__ void __launch_bounds__(kLnThreads)
layer_norm_bwd_kernel(const Y* __restrict__ dy, const X* __restrict__ x, const float* __restrict__ gamma,
                      const float* __restrict__ mean, const float* __restrict__ rstd, int64_t n, int c, X* __restrict__ dx,
                      float* __restrict__ part_g, float* __restrict__ part_b) {
  extern __shared__ float red[];         // [2][warps * rpw][c]
  const int L = c / (4 * V);
  const int rpw = 32 / L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % L, rin = lane / L;
  const int64_t warp_global = (blockIdx.x * (int64_t)kLnThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kLnThreads) >> 5;
  float4 g[V], ag[V], ab[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    g[i] = ld4<float>(gamma + 4 * (sub + i * L));
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.f / c;
  for (int64_t r0 = warp_global * rpw; r0 < n; r0 += n_warps * rpw) {
    const int64_t r = r0 + rin;
    const bool ok = r < n;
    const float mu = ok ? mean[r] : 0.f, rs = ok ? rstd[r] : 0.f;
    float4 xh[V], dh[V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 d = ok ? ld4<Y>(dy + r * c + 4 * (sub + i * L)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 xv = ok ? ld4<X>(x + r * c + 4 * (sub + i * L)) : make_float4(0.f, 0.f, 0.f, 0.f);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dh[i] = make_float4(d.x * g[i].x, d.y * g[i].y, d.z * g[i].z, d.w * g[i].w);
      s1 += (dh[i].x + dh[i].y) + (dh[i].z + dh[i].w);
      s2 += (dh[i].x * xh[i].x + dh[i].y * xh[i].y) + (dh[i].z * xh[i].z + dh[i].w * xh[i].w);
      ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
    }
    for (int o = L >> 1; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
    const float m1 = s1 * inv_c, m2 = s2 * inv_c;
    if (ok) {
#pragma unroll
      for (int i = 0; i < V; ++i)
        st4<X>(dx + r * c + 4 * (sub + i * L),
               make_float4(rs * (dh[i].x - m1 - xh[i].x * m2), rs * (dh[i].y - m1 - xh[i].y * m2),
                           rs * (dh[i].z - m1 - xh[i].z * m2), rs * (dh[i].w - m1 - xh[i].w * m2)));
    }
  }
  // block partials, fixed order: every (warp, row-in-warp) slot writes its channel vector, then the first c threads sum
  const int slots = (kLnThreads / 32) * rpw;
  float* rg = red;
  float* rb = red + slots * c;
  const int slot = warp * rpw + rin;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    *reinterpret_cast<float4*>(rg + slot * c + 4 * (sub + i * L)) = ag[i];
    *reinterpret_cast<float4*>(rb + slot * c + 4 * (sub + i * L)) = ab[i];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += kLnThreads) {
    float tg = 0.f, tb = 0.f;
    for (int sl = 0; sl < slots; ++sl) { tg += rg[sl * c + ch]; tb += rb[sl * c + ch]; }
    part_g[(int64_t)blockIdx.x * c + ch] = tg;
    part_b[(int64_t)blockIdx.x * c + ch] = tb;
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
  if (b > kNumSMs * 2) b = kNumSMs * 2;
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
  B2PC_CHECK_ARG(c == 32 || c == 64 || c == 128 || c == 256 || c == 512, "layer_norm: channels %d not one of 32/64/128/256/512", c);
  if (n == 0) return B2PC_OK;
  const int blocks = ln_blocks(n);
  if (c <= 128) { B2PC_LN_DISPATCH(xd, yd, (layer_norm_fwd_kernel<X, Y, 1><<<blocks, kLnThreads, 0, stream>>>((const X*)x, gamma, beta, n, c, eps, (Y*)y, mean, rstd))); }
  else if (c == 256) { B2PC_LN_DISPATCH(xd, yd, (layer_norm_fwd_kernel<X, Y, 2><<<blocks, kLnThreads, 0, stream>>>((const X*)x, gamma, beta, n, c, eps, (Y*)y, mean, rstd))); }
  else { B2PC_LN_DISPATCH(xd, yd, (layer_norm_fwd_kernel<X, Y, 4><<<blocks, kLnThreads, 0, stream>>>((const X*)x, gamma, beta, n, c, eps, (Y*)y, mean, rstd))); }
  count_launches(1);
  B2PC_CHECK_LAUNCH("layer_norm_fwd");
  return B2PC_OK;
}

inline int launch_layer_norm_bwd(const void* dy, int yd, const void* x, int xd, const float* gamma, const float* mean,
                                 const float* rstd, int64_t n, int c, void* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                                 cudaStream_t stream) {
  B2PC_CHECK_ARG(c == 32 || c == 64 || c == 128 || c == 256 || c == 512, "layer_norm: channels %d not one of 32/64/128/256/512", c);
  if (ws_bytes < layer_norm_bwd_workspace_bytes(n, c)) { set_error("layer_norm_bwd: workspace too small"); return B2PC_ERR_WORKSPACE; }
  const int blocks = ln_blocks(n);
  float* pg = (float*)ws;
  float* pb = pg + (int64_t)blocks * c;
  if (n == 0) {
    cudaMemsetAsync(dgamma, 0, c * sizeof(float), stream);
    if (dbeta) cudaMemsetAsync(dbeta, 0, c * sizeof(float), stream);
    return B2PC_OK;
  }
  const int vv = c <= 128 ? 1 : (c == 256 ? 2 : 4);
  const int rpw = 32 / (c / (4 * vv));
  const size_t smem = (size_t)2 * (kLnThreads / 32) * rpw * c * sizeof(float);   // <= 32 KB
  if (vv == 1) { B2PC_LN_DISPATCH(xd, yd, (layer_norm_bwd_kernel<X, Y, 1><<<blocks, kLnThreads, smem, stream>>>((const Y*)dy, (const X*)x, gamma, mean, rstd, n, c, (X*)dx, pg, pb))); }
  else if (vv == 2) { B2PC_LN_DISPATCH(xd, yd, (layer_norm_bwd_kernel<X, Y, 2><<<blocks, kLnThreads, smem, stream>>>((const Y*)dy, (const X*)x, gamma, mean, rstd, n, c, (X*)dx, pg, pb))); }
  else { B2PC_LN_DISPATCH(xd, yd, (layer_norm_bwd_kernel<X, Y, 4><<<blocks, kLnThreads, smem, stream>>>((const Y*)dy, (const X*)x, gamma, mean, rstd, n, c, (X*)dx, pg, pb))); }
  layer_norm_param_reduce_kernel<<<c / 32, 256, 0, stream>>>(pg, pb, blocks, c, dgamma, dbeta);
  count_launches(2);
  B2PC_CHECK_LAUNCH("layer_norm_bwd");
  return B2PC_OK;
}

}  // namespace b2pc

namespace b2pc {
// ---- column sum of a tall matrix: out[c] = sum_r x[r, c]  (bias gradients of the linears; fp32 result) ---------------------
// grid.x covers 256-channel slabs, grid.y row chunks; a block is 64 vector-columns x 4 row lanes, fully coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, int64_t n, int c, int64_t rows_per_block, float* __restrict__ partial) {
  __shared__ float4 red[4][64];
  const int cvi = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int ch = blockIdx.x * 256 + cvi * 4;
  const int64_t r_begin = blockIdx.y * rows_per_block;
  int64_t r_end = r_begin + rows_per_block;
  if (r_end > n) r_end = n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ch < c) {
    int64_t r = r_begin + rl;
    for (; r + 12 < r_end; r += 16) {   // 4 independent loads in flight
      const float4 a = ld4<T>(x + r * c + ch), b = ld4<T>(x + (r + 4) * c + ch), d = ld4<T>(x + (r + 8) * c + ch),
                   e = ld4<T>(x + (r + 12) * c + ch);
      acc.x += (a.x + b.x) + (d.x + e.x); acc.y += (a.y + b.y) + (d.y + e.y);
      acc.z += (a.z + b.z) + (d.z + e.z); acc.w += (a.w + b.w) + (d.w + e.w);
    }
    for (; r < r_end; r += 4) {
      const float4 a = ld4<T>(x + r * c + ch);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  red[rl][cvi] = acc;
  __syncthreads();
  if (rl == 0 && ch < c) {
    float4 t = red[0][cvi];
#pragma unroll
    for (int i = 1; i < 4; ++i) { t.x += red[i][cvi].x; t.y += red[i][cvi].y; t.z += red[i][cvi].z; t.w += red[i][cvi].w; }
    *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.y * c + ch) = t;
  }
}

// one block per 32 channels: 8 warps stride over the chunk partials, fixed-order shared-memory sum at the end
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int chunks, int c, float* __restrict__ out) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  float a = 0.f;
  if (ch < c)
    for (int b = grp; b < chunks; b += 8) a += partial[(int64_t)b * c + ch];
  sm[grp][lane] = a;
  __syncthreads();
  if (grp == 0 && ch < c) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    out[ch] = t;
  }
}

inline int colsum_chunks(int64_t n, int c) {
  const int slabs = (c + 255) / 256;
  int64_t chunks = (kNumSMs + slabs - 1) / slabs;
  const int64_t max_chunks = ceil_div(n > 0 ? n : 1, 64);
  if (chunks > max_chunks) chunks = max_chunks;
  return (int)(chunks < 1 ? 1 : chunks);
}
inline size_t colsum_workspace_bytes(int64_t n, int c) { return (size_t)colsum_chunks(n, c) * c * sizeof(float) + 256; }

inline int launch_colsum(const void* x, int dtype, int64_t n, int c, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  B2PC_CHECK_ARG(c % 4 == 0 && c > 0, "colsum: channels %d not a multiple of 4", c);
  if (ws_bytes < colsum_workspace_bytes(n, c)) { set_error("colsum: workspace too small"); return B2PC_ERR_WORKSPACE; }
  if (n == 0) { cudaMemsetAsync(out, 0, c * sizeof(float), stream); return B2PC_OK; }
  const int chunks = colsum_chunks(n, c);
  const int64_t rpb = ceil_div(ceil_div(n, chunks), 4) * 4;
  dim3 grid((c + 255) / 256, chunks);
  switch (dtype) {
    case B2PC_F32: colsum_partial_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, n, c, rpb, (float*)ws); break;
    case B2PC_F16: colsum_partial_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)x, n, c, rpb, (float*)ws); break;
    case B2PC_BF16: colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, n, c, rpb, (float*)ws); break;
    default: set_error("colsum: unknown dtype %d", dtype); return B2PC_ERR_INVALID_ARG;
  }
  colsum_final_kernel<<<(c + 31) / 32, 256, 0, stream>>>((const float*)ws, chunks, c, out);
  count_launches(2);
  B2PC_CHECK_LAUNCH("colsum");
  return B2PC_OK;
}
}  // namespace b2pc

namespace b2pc {
// ---- per-row scaling fused with the residual add (stochastic depth: out = shortcut + x * rowscale[row]) -------------------------
template <typename S, typename X>
__global__ void __launch_bounds__(256)
rowscale_add_kernel(const S* __restrict__ shortcut, const X* __restrict__ x, const float* __restrict__ rowscale, int64_t n, int c,
                    S* __restrict__ out) {
  const int64_t total = n * c / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = rowscale[(i * 4) / c];
    const float4 a = ld4<S>(shortcut + i * 4), b = ld4<X>(x + i * 4);
    st4<S>(out + i * 4, make_float4(fmaf(b.x, s, a.x), fmaf(b.y, s, a.y), fmaf(b.z, s, a.z), fmaf(b.w, s, a.w)));
  }
}
// dx = dy * rowscale[row]
template <typename Y, typename X>
__global__ void __launch_bounds__(256)
rowscale_kernel(const Y* __restrict__ dy, const float* __restrict__ rowscale, int64_t n, int c, X* __restrict__ dx) {
  const int64_t total = n * c / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = rowscale[(i * 4) / c];
    const float4 a = ld4<Y>(dy + i * 4);
    st4<X>(dx + i * 4, make_float4(a.x * s, a.y * s, a.z * s, a.w * s));
  }
}

#define B2PC_RS_DISPATCH(SD, XD, CALL)                                                              \
  if (SD == B2PC_F32 && XD == B2PC_F32) { using S = float; using X = float; CALL; }                 \
  else if (SD == B2PC_F32 && XD == B2PC_BF16) { using S = float; using X = __nv_bfloat16; CALL; }   \
  else if (SD == B2PC_F32 && XD == B2PC_F16) { using S = float; using X = __half; CALL; }           \
  else if (SD == B2PC_BF16 && XD == B2PC_BF16) { using S = __nv_bfloat16; using X = __nv_bfloat16; CALL; } \
  else if (SD == B2PC_F16 && XD == B2PC_F16) { using S = __half; using X = __half; CALL; }          \
  else { set_error("rowscale: unsupported dtype pair (%d, %d)", SD, XD); return B2PC_ERR_UNSUPPORTED; }

inline int launch_rowscale_add(const void* shortcut, int sd, const void* x, int xd, const float* rowscale, int64_t n, int c, void* out,
                               cudaStream_t stream) {
  B2PC_CHECK_ARG(c % 4 == 0 && c > 0, "rowscale_add: channels %d not a multiple of 4", c);
  if (n == 0) return B2PC_OK;
  int64_t b = ceil_div(n * c / 4, 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  B2PC_RS_DISPATCH(sd, xd, (rowscale_add_kernel<S, X><<<(int)b, 256, 0, stream>>>((const S*)shortcut, (const X*)x, rowscale, n, c, (S*)out)));
  count_launches(1);
  B2PC_CHECK_LAUNCH("rowscale_add");
  return B2PC_OK;
}
// dy in the shortcut/out dtype (sd), dx in x's dtype (xd)
inline int launch_rowscale(const void* dy, int sd, const float* rowscale, int64_t n, int c, void* dx, int xd, cudaStream_t stream) {
  B2PC_CHECK_ARG(c % 4 == 0 && c > 0, "rowscale: channels %d not a multiple of 4", c);
  if (n == 0) return B2PC_OK;
  int64_t b = ceil_div(n * c / 4, 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  B2PC_RS_DISPATCH(sd, xd, (rowscale_kernel<S, X><<<(int)b, 256, 0, stream>>>((const S*)dy, rowscale, n, c, (X*)dx)));
  count_launches(1);
  B2PC_CHECK_LAUNCH("rowscale");
  return B2PC_OK;
}
}  // namespace b2pc
