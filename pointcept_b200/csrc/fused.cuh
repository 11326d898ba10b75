// Fused residual glue of a PT-v3 block: the elementwise / row-normalisation work between two GEMM-shaped
// operators collapsed into ONE kernel per direction:
//
//     t = x                                  x: output of a Linear (half precision under autocast)
//     t = LayerNorm_a(t)                     optional  (the CPE's norm, ptv3m1:285)
//     t = t * dropscale(row)                 optional  (DropPath, ptv3m1:313-334: u[row] < keep ? 1/keep : 0)
//     r = shortcut + t                       fp32 residual stream                       -> written
//     r16 = half(r)                          optional  (input of the next block's sparse conv under autocast) -> written
//     y = LayerNorm_b(r)                     optional  (pre-norm of attention / MLP, emitted in the Linear's compute dtype) -> written
//
// which replaces up to five launches (LN, add, LN, bernoulli/div, cast) and their HBM round trips.  The backward is the exact
// adjoint, also one kernel (+ a tiny fixed-order reduction of the LayerNorm parameter gradients).
// HBM-bound: forward reads N*C*(4 + s) bytes and writes N*C*(4 + s [+ s]); rows are shared by L lanes as in layernorm.cuh.
#pragma once
#include "common.cuh"
#include "layernorm.cuh"

namespace b2pc {

struct FusedResArgs {
  const float* shortcut;    // [n, c] fp32
  const void* x;            // [n, c] T
  const float* u;           // [n] uniform randoms or null
  float keep;               // keep probability (only with u)
  const float* ga; const float* ba;   // LayerNorm_a affine (null = no LayerNorm_a); ba may be null
  const float* gb; const float* bb;   // LayerNorm_b affine (null = no LayerNorm_b)
  float eps_a, eps_b;
  int64_t n; int c;
  float* r;                 // [n, c] fp32 out
  void* r16;                // [n, c] T out or null
  void* y;                  // [n, c] T out or null (requires gb)
  float* stat_a;            // [2, n] mean / rstd of LayerNorm_a
  float* stat_b;            // [2, n]
};

template <typename T, int V>
__global__ void __launch_bounds__(kLnThreads)
fused_residual_fwd_kernel(FusedResArgs a) {
  const int c = a.c;
  const int64_t n = a.n;
  const int L = c / (4 * V);
  const int rpw = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, rin = lane / L;
  const int64_t warp_global = (blockIdx.x * (int64_t)kLnThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kLnThreads) >> 5;
  const T* x = reinterpret_cast<const T*>(a.x);
  T* r16 = reinterpret_cast<T*>(a.r16);
  T* y = reinterpret_cast<T*>(a.y);
  const bool ln_a = a.ga != nullptr, ln_b = a.gb != nullptr;
  float4 ga[V], ba[V], gb[V], bb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int o = 4 * (sub + i * L);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    ga[i] = ln_a ? ld4<float>(a.ga + o) : z;
    ba[i] = (ln_a && a.ba) ? ld4<float>(a.ba + o) : z;
    gb[i] = ln_b ? ld4<float>(a.gb + o) : z;
    bb[i] = (ln_b && a.bb) ? ld4<float>(a.bb + o) : z;
  }
  const float inv_c = 1.f / c;
  const float inv_keep = a.u ? 1.f / a.keep : 1.f;
  for (int64_t r0 = warp_global * rpw; r0 < n; r0 += n_warps * rpw) {
    const int64_t row = r0 + rin;
    const bool ok = row < n;
    float4 v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = ok ? ld4<T>(x + row * c + 4 * (sub + i * L)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (ln_a) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      const float mu = s * inv_c;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
      for (int o = L >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
      const float rs = rsqrtf(q * inv_c + a.eps_a);
#pragma unroll
      for (int i = 0; i < V; ++i)
        v[i] = make_float4((v[i].x - mu) * rs * ga[i].x + ba[i].x, (v[i].y - mu) * rs * ga[i].y + ba[i].y,
                           (v[i].z - mu) * rs * ga[i].z + ba[i].z, (v[i].w - mu) * rs * ga[i].w + ba[i].w);
      if (ok && sub == 0) { a.stat_a[row] = mu; a.stat_a[n + row] = rs; }
    }
    float sc = 1.f;
    if (a.u && ok) sc = a.u[row] < a.keep ? inv_keep : 0.f;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 h = ok ? ld4<float>(a.shortcut + row * c + 4 * (sub + i * L)) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[i] = make_float4(fmaf(v[i].x, sc, h.x), fmaf(v[i].y, sc, h.y), fmaf(v[i].z, sc, h.z), fmaf(v[i].w, sc, h.w));
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    if (ok) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        st4<float>(a.r + row * c + 4 * (sub + i * L), v[i]);
        if (r16) st4<T>(r16 + row * c + 4 * (sub + i * L), v[i]);
      }
    }
    if (ln_b) {
      for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      const float mu = s * inv_c;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
      for (int o = L >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
      const float rs = rsqrtf(q * inv_c + a.eps_b);
      if (ok) {
#pragma unroll
        for (int i = 0; i < V; ++i)
          st4<T>(y + row * c + 4 * (sub + i * L),
                 make_float4((v[i].x - mu) * rs * gb[i].x + bb[i].x, (v[i].y - mu) * rs * gb[i].y + bb[i].y,
                             (v[i].z - mu) * rs * gb[i].z + bb[i].z, (v[i].w - mu) * rs * gb[i].w + bb[i].w));
        if (sub == 0) { a.stat_b[row] = mu; a.stat_b[n + row] = rs; }
      }
    }
  }
}

struct FusedResBwdArgs {
  const float* dr_out;      // [n, c] fp32 gradient wrt r (null = none)
  const void* dr16;         // [n, c] T gradient wrt r16 (null = none)
  const void* dy;           // [n, c] T gradient wrt y (null = none)
  const float* r;           // saved r (needed with LayerNorm_b)
  const void* x;            // saved x (needed with LayerNorm_a)
  const float* u; float keep;
  const float* ga; const float* gb;
  const float* stat_a; const float* stat_b;
  int64_t n; int c;
  float* d_shortcut;        // [n, c] fp32 out
  void* dx;                 // [n, c] T out
  float* part;              // [blocks][5][c]: dga, dba, dgb, dbb, colsum(dx) partial sums
  int want_dx_colsum;       // also reduce the column sums of dx (= bias gradient of the Linear that produced x)
};

template <typename T, int V>
__global__ void __launch_bounds__(kLnThreads)
fused_residual_bwd_kernel(FusedResBwdArgs a) {
  extern __shared__ float red[];         // [2][slots][c]
  const int c = a.c;
  const int64_t n = a.n;
  const int L = c / (4 * V);
  const int rpw = 32 / L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % L, rin = lane / L;
  const int64_t warp_global = (blockIdx.x * (int64_t)kLnThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kLnThreads) >> 5;
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* dy = reinterpret_cast<const T*>(a.dy);
  const T* dr16 = reinterpret_cast<const T*>(a.dr16);
  T* dx = reinterpret_cast<T*>(a.dx);
  const bool ln_a = a.ga != nullptr, ln_b = a.gb != nullptr && a.dy != nullptr;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ga[V], gb[V], aga[V], aba[V], agb[V], abb[V], adx[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int o = 4 * (sub + i * L);
    ga[i] = ln_a ? ld4<float>(a.ga + o) : z4;
    gb[i] = ln_b ? ld4<float>(a.gb + o) : z4;
    aga[i] = aba[i] = agb[i] = abb[i] = adx[i] = z4;
  }
  const float inv_c = 1.f / c;
  const float inv_keep = a.u ? 1.f / a.keep : 1.f;
  for (int64_t r0 = warp_global * rpw; r0 < n; r0 += n_warps * rpw) {
    const int64_t row = r0 + rin;
    const bool ok = row < n;
    float4 dr[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int64_t o = row * c + 4 * (sub + i * L);
      dr[i] = (ok && a.dr_out) ? ld4<float>(a.dr_out + o) : z4;
      if (ok && dr16) { const float4 h = ld4<T>(dr16 + o); dr[i].x += h.x; dr[i].y += h.y; dr[i].z += h.z; dr[i].w += h.w; }
    }
    if (ln_b) {
      const float mu = ok ? a.stat_b[row] : 0.f, rs = ok ? a.stat_b[n + row] : 0.f;
      float4 xh[V], dh[V];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int64_t o = row * c + 4 * (sub + i * L);
        const float4 d = ok ? ld4<T>(dy + o) : z4;
        const float4 rv = ok ? ld4<float>(a.r + o) : z4;
        xh[i] = make_float4((rv.x - mu) * rs, (rv.y - mu) * rs, (rv.z - mu) * rs, (rv.w - mu) * rs);
        dh[i] = make_float4(d.x * gb[i].x, d.y * gb[i].y, d.z * gb[i].z, d.w * gb[i].w);
        s1 += (dh[i].x + dh[i].y) + (dh[i].z + dh[i].w);
        s2 += (dh[i].x * xh[i].x + dh[i].y * xh[i].y) + (dh[i].z * xh[i].z + dh[i].w * xh[i].w);
        agb[i].x += d.x * xh[i].x; agb[i].y += d.y * xh[i].y; agb[i].z += d.z * xh[i].z; agb[i].w += d.w * xh[i].w;
        abb[i].x += d.x; abb[i].y += d.y; abb[i].z += d.z; abb[i].w += d.w;
      }
      for (int o = L >> 1; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
      const float m1 = s1 * inv_c, m2 = s2 * inv_c;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        dr[i].x += rs * (dh[i].x - m1 - xh[i].x * m2); dr[i].y += rs * (dh[i].y - m1 - xh[i].y * m2);
        dr[i].z += rs * (dh[i].z - m1 - xh[i].z * m2); dr[i].w += rs * (dh[i].w - m1 - xh[i].w * m2);
      }
    }
    float sc = 1.f;
    if (a.u && ok) sc = a.u[row] < a.keep ? inv_keep : 0.f;
    if (ok) {
#pragma unroll
      for (int i = 0; i < V; ++i) st4<float>(a.d_shortcut + row * c + 4 * (sub + i * L), dr[i]);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { dr[i].x *= sc; dr[i].y *= sc; dr[i].z *= sc; dr[i].w *= sc; }   // dr is now dt
    if (ln_a) {
      const float mu = ok ? a.stat_a[row] : 0.f, rs = ok ? a.stat_a[n + row] : 0.f;
      float4 xh[V], dh[V];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 xv = ok ? ld4<T>(x + row * c + 4 * (sub + i * L)) : z4;
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        dh[i] = make_float4(dr[i].x * ga[i].x, dr[i].y * ga[i].y, dr[i].z * ga[i].z, dr[i].w * ga[i].w);
        s1 += (dh[i].x + dh[i].y) + (dh[i].z + dh[i].w);
        s2 += (dh[i].x * xh[i].x + dh[i].y * xh[i].y) + (dh[i].z * xh[i].z + dh[i].w * xh[i].w);
        aga[i].x += dr[i].x * xh[i].x; aga[i].y += dr[i].y * xh[i].y; aga[i].z += dr[i].z * xh[i].z; aga[i].w += dr[i].w * xh[i].w;
        aba[i].x += dr[i].x; aba[i].y += dr[i].y; aba[i].z += dr[i].z; aba[i].w += dr[i].w;
      }
      for (int o = L >> 1; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
      const float m1 = s1 * inv_c, m2 = s2 * inv_c;
      if (ok) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float4 o4 = make_float4(rs * (dh[i].x - m1 - xh[i].x * m2), rs * (dh[i].y - m1 - xh[i].y * m2),
                                        rs * (dh[i].z - m1 - xh[i].z * m2), rs * (dh[i].w - m1 - xh[i].w * m2));
          st4<T>(dx + row * c + 4 * (sub + i * L), o4);
          adx[i].x += o4.x; adx[i].y += o4.y; adx[i].z += o4.z; adx[i].w += o4.w;
        }
      }
    } else if (ok) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        st4<T>(dx + row * c + 4 * (sub + i * L), dr[i]);
        adx[i].x += dr[i].x; adx[i].y += dr[i].y; adx[i].z += dr[i].z; adx[i].w += dr[i].w;
      }
    }
  }
  // block partials of the four parameter gradients, two at a time through shared memory, fixed summation order
  const int slots = (kLnThreads / 32) * rpw;
  const int slot = warp * rpw + rin;
  float* r0s = red;
  float* r1s = red + slots * c;
  for (int pass = 0; pass < 3; ++pass) {
    if ((pass == 0 && !ln_a) || (pass == 1 && !ln_b) || (pass == 2 && !a.want_dx_colsum)) continue;   // uniform across the grid
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
      *reinterpret_cast<float4*>(r0s + slot * c + 4 * (sub + i * L)) = pass == 0 ? aga[i] : (pass == 1 ? agb[i] : adx[i]);
      *reinterpret_cast<float4*>(r1s + slot * c + 4 * (sub + i * L)) = pass == 0 ? aba[i] : abb[i];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += kLnThreads) {
      float t0 = 0.f, t1 = 0.f;
      for (int sl = 0; sl < slots; ++sl) { t0 += r0s[sl * c + ch]; t1 += r1s[sl * c + ch]; }
      if (pass < 2) {
        a.part[((int64_t)blockIdx.x * 5 + 2 * pass) * c + ch] = t0;
        a.part[((int64_t)blockIdx.x * 5 + 2 * pass + 1) * c + ch] = t1;
      } else {
        a.part[((int64_t)blockIdx.x * 5 + 4) * c + ch] = t0;
      }
    }
  }
}

// out[v][ch] = sum_b part[b][v][ch] for the parameter vectors v whose destination pointer is non-null; block = 32 channels
__global__ void __launch_bounds__(256)
fused_param_reduce_kernel(const float* __restrict__ part, int blocks, int c, float* __restrict__ o0, float* __restrict__ o1,
                          float* __restrict__ o2, float* __restrict__ o3, float* __restrict__ o4) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  float* outs[5] = {o0, o1, o2, o3, o4};
  const int v = blockIdx.y;                 // one parameter vector per grid row
  if (!outs[v]) return;
  float acc0 = 0.f, acc1 = 0.f;             // two independent chains (fixed order: deterministic)
  int b = grp;
  for (; b + 8 < blocks; b += 16) {
    acc0 += part[((int64_t)b * 5 + v) * c + ch];
    acc1 += part[((int64_t)(b + 8) * 5 + v) * c + ch];
  }
  if (b < blocks) acc0 += part[((int64_t)b * 5 + v) * c + ch];
  sm[grp][lane] = acc0 + acc1;
  __syncthreads();
  if (grp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    outs[v][ch] = t;
  }
}

inline bool fused_channels_ok(int c) { return c == 32 || c == 64 || c == 128 || c == 256 || c == 512; }
// HBM-bound streaming kernels: enough resident warps to keep tens of KB in flight per SM (ncu: 2 blocks per SM reached 27 % of
// the copy bandwidth); the backward is limited to 4 blocks per SM by its 32 KB reduction scratch
inline int fused_blocks(int64_t n, int per_sm) {
  int64_t b = ceil_div(n, kLnThreads / 32 * 4);
  if (b > kNumSMs * per_sm) b = kNumSMs * per_sm;
  return (int)(b < 1 ? 1 : b);
}
inline size_t fused_residual_bwd_workspace_bytes(int64_t n, int c) { return (size_t)fused_blocks(n, 4) * 5 * c * sizeof(float) + 256; }

#define B2PC_FR_DISPATCH(DT, VV, KERNEL, ...)                                                          \
  do {                                                                                                 \
    if (DT == B2PC_F32) { if (VV == 1) KERNEL<float, 1> __VA_ARGS__; else if (VV == 2) KERNEL<float, 2> __VA_ARGS__; else KERNEL<float, 4> __VA_ARGS__; }  \
    else if (DT == B2PC_BF16) { if (VV == 1) KERNEL<__nv_bfloat16, 1> __VA_ARGS__; else if (VV == 2) KERNEL<__nv_bfloat16, 2> __VA_ARGS__; else KERNEL<__nv_bfloat16, 4> __VA_ARGS__; } \
    else { if (VV == 1) KERNEL<__half, 1> __VA_ARGS__; else if (VV == 2) KERNEL<__half, 2> __VA_ARGS__; else KERNEL<__half, 4> __VA_ARGS__; } \
  } while (0)

inline int launch_fused_residual_fwd(const FusedResArgs& a, int dtype, cudaStream_t stream) {
  B2PC_CHECK_ARG(fused_channels_ok(a.c), "fused_residual: channels %d not one of 32/64/128/256/512", a.c);
  B2PC_CHECK_ARG(dtype == B2PC_F32 || dtype == B2PC_F16 || dtype == B2PC_BF16, "fused_residual: bad dtype %d", dtype);
  B2PC_CHECK_ARG(a.shortcut && a.x && a.r, "fused_residual_fwd: null pointer");
  B2PC_CHECK_ARG(!a.y || a.gb, "fused_residual_fwd: y requested without LayerNorm_b parameters");
  B2PC_CHECK_ARG((!a.ga || a.stat_a) && (!a.gb || a.stat_b), "fused_residual_fwd: statistics buffer missing");
  if (a.n == 0) return B2PC_OK;
  const int vv = a.c <= 128 ? 1 : (a.c == 256 ? 2 : 4);
  const int blocks = fused_blocks(a.n, 8);
  B2PC_FR_DISPATCH(dtype, vv, fused_residual_fwd_kernel, <<<blocks, kLnThreads, 0, stream>>>(a));
  count_launches(1);
  B2PC_CHECK_LAUNCH("fused_residual_fwd");
  return B2PC_OK;
}

inline int launch_fused_residual_bwd(const FusedResBwdArgs& a0, int dtype, float* dga, float* dba, float* dgb, float* dbb, float* dx_colsum,
                                     void* ws, size_t ws_bytes, cudaStream_t stream) {
  FusedResBwdArgs a = a0;
  a.want_dx_colsum = dx_colsum != nullptr;
  B2PC_CHECK_ARG(fused_channels_ok(a.c), "fused_residual: channels %d not one of 32/64/128/256/512", a.c);
  B2PC_CHECK_ARG(dtype == B2PC_F32 || dtype == B2PC_F16 || dtype == B2PC_BF16, "fused_residual: bad dtype %d", dtype);
  B2PC_CHECK_ARG(a.d_shortcut && a.dx && ws, "fused_residual_bwd: null pointer");
  if (ws_bytes < fused_residual_bwd_workspace_bytes(a.n, a.c)) { set_error("fused_residual_bwd: workspace too small"); return B2PC_ERR_WORKSPACE; }
  const bool ln_a = a.ga != nullptr, ln_b = a.gb != nullptr && a.dy != nullptr;
  B2PC_CHECK_ARG((!ln_a || (a.x && a.stat_a && dga)) && (!ln_b || (a.r && a.stat_b && dgb)), "fused_residual_bwd: saved tensors missing");
  if (a.gb && !a.dy) {   // LayerNorm_b output unused downstream: its parameters get zero gradients
    if (dgb) cudaMemsetAsync(dgb, 0, a.c * sizeof(float), stream);
    if (dbb) cudaMemsetAsync(dbb, 0, a.c * sizeof(float), stream);
  }
  if (a.n == 0) {
    float* outs[5] = {ln_a ? dga : nullptr, ln_a ? dba : nullptr, ln_b ? dgb : nullptr, ln_b ? dbb : nullptr, dx_colsum};
    for (float* o : outs) if (o) cudaMemsetAsync(o, 0, a.c * sizeof(float), stream);
    return B2PC_OK;
  }
  a.part = (float*)ws;
  const int vv = a.c <= 128 ? 1 : (a.c == 256 ? 2 : 4);
  const int rpw = 32 / (a.c / (4 * vv));
  const size_t smem = (size_t)2 * (kLnThreads / 32) * rpw * a.c * sizeof(float);   // <= 32 KB
  const int blocks = fused_blocks(a.n, 4);
  B2PC_FR_DISPATCH(dtype, vv, fused_residual_bwd_kernel, <<<blocks, kLnThreads, smem, stream>>>(a));
  count_launches(1);
  if (ln_a || ln_b || dx_colsum) {
    fused_param_reduce_kernel<<<dim3(a.c / 32, 5), 256, 0, stream>>>(a.part, blocks, a.c, ln_a ? dga : nullptr, ln_a ? dba : nullptr,
                                                                     ln_b ? dgb : nullptr, ln_b ? dbb : nullptr, dx_colsum);
    count_launches(1);
  }
  B2PC_CHECK_LAUNCH("fused_residual_bwd");
  return B2PC_OK;
}

// ---- one launch casts a whole list of fp32 parameter tensors to their half-precision shadows --------------------------------
// plan[i] = {src pointer, dst pointer, element count, first block}; built once by the host binding (parameter storage is stable)
struct CastItem { const float* src; void* dst; long long count; long long first_block; };
constexpr int kCastBlockElems = 256 * 8;

template <typename T>
__global__ void __launch_bounds__(256)
multi_cast_kernel(const CastItem* __restrict__ plan, int n_items) {
  // binary search for the item owning this block
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (plan[mid].first_block <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const CastItem it = plan[lo];
  const long long base = ((long long)blockIdx.x - it.first_block) * kCastBlockElems;
  T* dst = reinterpret_cast<T*>(it.dst);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const long long i = base + e * 256 + threadIdx.x;
    if (i < it.count) dst[i] = from_f32<T>(it.src[i]);
  }
}

inline int launch_multi_cast(const void* plan, int n_items, long long total_blocks, int dst_dtype, cudaStream_t stream) {
  B2PC_CHECK_ARG(plan && n_items >= 0 && total_blocks >= 0, "multi_cast: bad arguments");
  B2PC_CHECK_ARG(dst_dtype == B2PC_F16 || dst_dtype == B2PC_BF16 || dst_dtype == B2PC_F32, "multi_cast: unknown destination dtype %d", dst_dtype);
  if (n_items == 0 || total_blocks == 0) return B2PC_OK;
  if (dst_dtype == B2PC_BF16) multi_cast_kernel<__nv_bfloat16><<<(unsigned)total_blocks, 256, 0, stream>>>((const CastItem*)plan, n_items);
  else if (dst_dtype == B2PC_F32)   // plain multi-tensor copy: packs a list of gradients into one flat all-reduce buffer (reducer.py)
    multi_cast_kernel<float><<<(unsigned)total_blocks, 256, 0, stream>>>((const CastItem*)plan, n_items);
  else multi_cast_kernel<__half><<<(unsigned)total_blocks, 256, 0, stream>>>((const CastItem*)plan, n_items);
  count_launches(1);
  B2PC_CHECK_LAUNCH("multi_cast");
  return B2PC_OK;
}

// ---- exact GELU (erf form, nn.GELU default as used at ptv3m1:233) forward / backward, one pass each ------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const T* __restrict__ x, int64_t total4, T* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld4<T>(x + i * 4);
    auto g = [](float t) { return 0.5f * t * (1.f + erff(t * 0.70710678118654752f)); };
    st4<T>(y + i * 4, make_float4(g(v.x), g(v.y), g(v.z), g(v.w)));
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, int64_t total4, T* __restrict__ dx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld4<T>(x + i * 4), d = ld4<T>(dy + i * 4);
    auto g = [](float t, float dd) {
      const float cdf = 0.5f * (1.f + erff(t * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * t * t);
      return dd * (cdf + t * pdf);
    };
    st4<T>(dx + i * 4, make_float4(g(v.x, d.x), g(v.y, d.y), g(v.z, d.z), g(v.w, d.w)));
  }
}


// GELU backward fused with the column sums of its result (= bias gradient of the Linear in front of the GELU, ptv3m1:238):
// same tiling as colsum_partial_kernel (256-channel slabs x row chunks, 64 vector columns x 4 row lanes per block).
template <typename T>
__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(const T* __restrict__ dy, const T* __restrict__ x, int64_t n, int c, int64_t rows_per_block, T* __restrict__ dx,
                       float* __restrict__ partial) {
  __shared__ float4 red[4][64];
  const int cvi = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int ch = blockIdx.x * 256 + cvi * 4;
  const int64_t r_begin = blockIdx.y * rows_per_block;
  int64_t r_end = r_begin + rows_per_block;
  if (r_end > n) r_end = n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto g = [](float t, float dd) {
    const float cdf = 0.5f * (1.f + erff(t * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * t * t);
    return dd * (cdf + t * pdf);
  };
  if (ch < c) {
    for (int64_t r = r_begin + rl; r < r_end; r += 4) {
      const float4 v = ld4<T>(x + r * c + ch), d = ld4<T>(dy + r * c + ch);
      const float4 o = make_float4(g(v.x, d.x), g(v.y, d.y), g(v.z, d.z), g(v.w, d.w));
      st4<T>(dx + r * c + ch, o);
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
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

inline int gelu_colsum_chunks(int64_t n, int c) {
  const int slabs = (c + 255) / 256;
  int64_t chunks = (4 * kNumSMs + slabs - 1) / slabs;
  const int64_t max_chunks = ceil_div(n > 0 ? n : 1, 32);
  if (chunks > max_chunks) chunks = max_chunks;
  return (int)(chunks < 1 ? 1 : chunks);
}
inline size_t gelu_bwd_colsum_workspace_bytes(int64_t n, int c) { return (size_t)gelu_colsum_chunks(n, c) * c * sizeof(float) + 256; }

inline int launch_gelu_bwd_colsum(const void* dy, const void* x, int dtype, int64_t n, int c, void* dx, float* colsum_out, void* ws,
                                  size_t ws_bytes, cudaStream_t stream) {
  B2PC_CHECK_ARG(c % 4 == 0 && c > 0 && n >= 0, "gelu_bwd_colsum: bad sizes");
  if (ws_bytes < gelu_bwd_colsum_workspace_bytes(n, c)) { set_error("gelu_bwd_colsum: workspace too small"); return B2PC_ERR_WORKSPACE; }
  if (n == 0) { cudaMemsetAsync(colsum_out, 0, c * sizeof(float), stream); return B2PC_OK; }
  const int chunks = gelu_colsum_chunks(n, c);
  const int64_t rpb = ceil_div(ceil_div(n, chunks), 4) * 4;
  dim3 grid((c + 255) / 256, chunks);
  switch (dtype) {
    case B2PC_F32: gelu_bwd_colsum_kernel<float><<<grid, 256, 0, stream>>>((const float*)dy, (const float*)x, n, c, rpb, (float*)dx, (float*)ws); break;
    case B2PC_F16: gelu_bwd_colsum_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)dy, (const __half*)x, n, c, rpb, (__half*)dx, (float*)ws); break;
    case B2PC_BF16: gelu_bwd_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, n, c, rpb, (__nv_bfloat16*)dx, (float*)ws); break;
    default: set_error("gelu_bwd_colsum: unknown dtype %d", dtype); return B2PC_ERR_INVALID_ARG;
  }
  colsum_final_kernel<<<(c + 31) / 32, 256, 0, stream>>>((const float*)ws, chunks, c, colsum_out);
  count_launches(2);
  B2PC_CHECK_LAUNCH("gelu_bwd_colsum");
  return B2PC_OK;
}


// ---- AdamW over a whole parameter list in ONE launch (SURVEY.md 8(f).2 "fused AdamW behind the all-reduce") ---------------------
// torch.optim.AdamW semantics (decoupled weight decay, bias correction), fp32 parameters / gradients / moments:
//   p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// items[i] = {p, g, m, v, count, first_block, bc1 = 1 - b1^t, sqrt(bc2 = 1 - b2^t)} with t the tensor's own step count;
// a block updates 2048 consecutive elements of one tensor.
struct AdamItem { float* p; const float* g; float* m; float* v; long long count; long long first_block; float bc1; float bc2_sqrt; };

__global__ void __launch_bounds__(256)
multi_adamw_kernel(const AdamItem* __restrict__ items, int n_items, float lr, float b1, float b2, float eps, float wd, float grad_scale) {
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const AdamItem it = items[lo];
  const long long base = ((long long)blockIdx.x - it.first_block) * kCastBlockElems;
  const float step = lr / it.bc1, decay = 1.f - lr * wd, bc2_sqrt = it.bc2_sqrt;   // per-tensor bias correction (own step count)
  const bool vec = ((reinterpret_cast<uintptr_t>(it.p) | reinterpret_cast<uintptr_t>(it.g) | reinterpret_cast<uintptr_t>(it.m) |
                     reinterpret_cast<uintptr_t>(it.v)) & 15) == 0;
  auto upd = [&](float& p, float g, float& m, float& v) {
    g *= grad_scale;
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    p = p * decay - step * m / (sqrtf(v) / bc2_sqrt + eps);
  };
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const long long i = base + (e * 256 + threadIdx.x) * 4;
    if (i >= it.count) continue;
    if (vec && i + 4 <= it.count) {
      float4 p = *reinterpret_cast<float4*>(it.p + i), m = *reinterpret_cast<float4*>(it.m + i), v = *reinterpret_cast<float4*>(it.v + i);
      const float4 g = *reinterpret_cast<const float4*>(it.g + i);
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(it.p + i) = p;
      *reinterpret_cast<float4*>(it.m + i) = m;
      *reinterpret_cast<float4*>(it.v + i) = v;
    } else {
      for (long long j = i; j < i + 4 && j < it.count; ++j) {
        float p = it.p[j], m = it.m[j], v = it.v[j];
        upd(p, it.g[j], m, v);
        it.p[j] = p; it.m[j] = m; it.v[j] = v;
      }
    }
  }
}

inline int launch_multi_adamw(const void* items, int n_items, long long total_blocks, float lr, float b1, float b2, float eps, float wd,
                              float grad_scale, cudaStream_t stream) {
  B2PC_CHECK_ARG(items && n_items >= 0 && total_blocks >= 0, "multi_adamw: bad arguments");
  if (n_items == 0 || total_blocks == 0) return B2PC_OK;
  multi_adamw_kernel<<<(unsigned)total_blocks, 256, 0, stream>>>((const AdamItem*)items, n_items, lr, b1, b2, eps, wd, grad_scale);
  count_launches(1);
  B2PC_CHECK_LAUNCH("multi_adamw");
  return B2PC_OK;
}

}  // namespace b2pc
