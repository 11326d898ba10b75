// Patch attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), head_dim 16.
//
// One CTA = one (sequence, head, 128-query tile); 128 threads, thread t owns query row t (= TMEM lane t), so the
// row max / row sum of the softmax need no cross-thread traffic.  Per block of BN keys:
//     S  = Q K_j^T        one tcgen05.mma M=128 N=BN K=16, operands from shared memory      -> TMEM columns [0,BN)
//     P  = exp2(c*S - m)  each thread reads its S row with tcgen05.ld, writes bf16/fp16 P    -> TMEM columns [BN,BN+BN/2)
//     PV = P V_j          BN/16 tcgen05.mma M=128 N=16 K=16, A from TMEM, B = V (MN-major)   -> TMEM columns [1.5BN, +16)
// and the running output row lives in registers (O = O*corr + PV).  K/V blocks stream through a 3-stage cp.async ring.
// D = 16 makes this kernel exp-bound (64 MMA-flop per exp), not tensor-bound; several CTAs per SM overlap the MMA latency
// of one tile with the softmax of another.
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace b2pc {

struct AttnDesc { uint32_t q_lbo, q_sbo, k_lbo, k_sbo, v_lbo, v_sbo; };

template <typename T> struct UmmaFmt;
template <> struct UmmaFmt<__nv_bfloat16> { static constexpr int v = umma::kFmtBF16; };
template <> struct UmmaFmt<__half> { static constexpr int v = umma::kFmtF16; };

template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int kAuQ = 128;      // queries per CTA
constexpr int kAuStages = 3;   // K/V ring depth

template <int BN> __host__ __device__ constexpr int attn_tmem_cols() { return BN == 128 ? 256 : 128; }
template <int BN> __host__ __device__ constexpr int attn_fwd_smem_bytes() { return kAuQ * 32 + kAuStages * BN * 64 + 64; }

template <typename T, int BN>
__global__ void __launch_bounds__(kAuQ)
attn_fwd_umma_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ cu, int64_t t_total, int H, float scale,
                     T* __restrict__ out, float* __restrict__ lse, AttnDesc dd) {
  using namespace umma;
  constexpr int D = 16;
  constexpr int TMEM_COLS = attn_tmem_cols<BN>();
  constexpr uint32_t COL_S = 0, COL_P = BN, COL_O = BN + BN / 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* q_s = smem;
  uint8_t* kv_s = smem + kAuQ * 32;
  uint64_t* bar = reinterpret_cast<uint64_t*>(kv_s + kAuStages * BN * 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int seq = blockIdx.y, h = blockIdx.z;
  const int64_t s0 = cu[seq];
  const int len = (int)(cu[seq + 1] - s0);
  const int q0 = blockIdx.x * kAuQ;
  if (q0 >= len) return;
  const int nblk = (len + BN - 1) / BN;
  const int64_t row_stride = (int64_t)3 * H * D;  // elements between consecutive tokens

  if (warp == 0) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }

  // ---- loads: 16-byte pieces into "plane" layout: piece (row r, chunk c) -> c * rows*16 + r*16 ----------------------
  const T* base_q = qkv + (s0 * 3 + 0) * H * D + h * D;
  const T* base_k = qkv + (s0 * 3 + 1) * H * D + h * D;
  const T* base_v = qkv + (s0 * 3 + 2) * H * D + h * D;
  {
    const int r = tid;
    const bool ok = q0 + r < len;
    const T* src = base_q + (int64_t)(q0 + r) * row_stride;
    cp_async16(smem_u32(q_s + r * 16), ok ? src : base_q, ok);
    cp_async16(smem_u32(q_s + kAuQ * 16 + r * 16), ok ? src + 8 : base_q, ok);
  }
  auto load_kv = [&](int blk, int stage) {
    uint8_t* ks = kv_s + stage * BN * 64;
    uint8_t* vs = ks + BN * 32;
    const int k0 = blk * BN;
    for (int p = tid; p < 4 * BN; p += kAuQ) {
      const int which = p / (2 * BN), rem = p % (2 * BN), r = rem >> 1, c = rem & 1;
      const bool ok = k0 + r < len;
      const T* src = (which ? base_v : base_k) + (int64_t)(k0 + r) * row_stride + c * 8;
      cp_async16(smem_u32((which ? vs : ks) + c * BN * 16 + r * 16), ok ? src : base_k, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();
  if (nblk > 1) load_kv(1, 1);
  cp_async_commit();

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);

  constexpr uint32_t idesc_s = make_idesc(128, BN, UmmaFmt<T>::v, UmmaFmt<T>::v, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc(128, 16, UmmaFmt<T>::v, UmmaFmt<T>::v, 0, 1);
  const uint64_t desc_q = make_smem_desc(smem_u32(q_s), dd.q_lbo, dd.q_sbo);

  const float c = scale * kLog2e;
  float m = -INFINITY, l = 0.f, o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;

  for (int j = 0; j < nblk; ++j) {
    const int stage = j % kAuStages;
    uint8_t* ks = kv_s + stage * BN * 64;
    uint8_t* vs = ks + BN * 32;
    cp_async_wait<1>();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_ss(tmem_base + COL_S, desc_q, make_smem_desc(smem_u32(ks), dd.k_lbo, dd.k_sbo), idesc_s, 0);
      mma_commit(bar);
    }
    mbar_wait(bar, j & 1);
    tc_fence_after();
    // S_j is complete, hence so is PV_{j-1}: its K/V stage is free again -> prefetch block j+2 into it
    if (j + 2 < nblk) load_kv(j + 2, (j + 2) % kAuStages);
    cp_async_commit();
    if (j > 0) {
      uint32_t r[16];
      tmem_ld16(lane_base + COL_O, r);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] += __uint_as_float(r[d]);
    }
    const int k0 = j * BN;
    const bool tail = (j == nblk - 1) && (len - k0 < BN);
    const int nvalid = len - k0;
    // pass A: row maximum of this block
    float mx = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t r[32];
      tmem_ld32(lane_base + COL_S + ch * 32, r);
      tmem_ld_wait();
      if (!tail) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (ch * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
    const float m_new = fmaxf(m, mx * c);
    const float corr = ex2(m - m_new);
    l *= corr;
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] *= corr;
    // pass B: P = exp2(c*S - m_new), row sum, pack to the MMA operand type, store to TMEM
    float lsum = 0.f;
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t r[32];
      tmem_ld32(lane_base + COL_S + ch * 32, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), c, -m_new));
        float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), c, -m_new));
        if (tail) {
          if (ch * 32 + 2 * i >= nvalid) p0 = 0.f;
          if (ch * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
        }
        lsum += p0 + p1;
        pk[i] = pack2<T>(p0, p1);
      }
      tmem_st16(lane_base + COL_P + ch * 16, pk);
    }
    l += lsum;
    m = m_new;
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < BN / 16; ++kk)
        mma_ts(tmem_base + COL_O, tmem_base + COL_P + kk * 8, make_smem_desc(smem_u32(vs + kk * 256), dd.v_lbo, dd.v_sbo),
               idesc_pv, kk > 0);
      if (j == nblk - 1) mma_commit(bar);
    }
  }
  mbar_wait(bar, nblk & 1);
  tc_fence_after();
  {
    uint32_t r[16];
    tmem_ld16(lane_base + COL_O, r);
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] += __uint_as_float(r[d]);
  }
  const int qi = q0 + tid;
  if (qi < len) {
    const float inv = 1.f / l;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack2<T>(o[2 * i] * inv, o[2 * i + 1] * inv);
    uint4* dst = reinterpret_cast<uint4*>(out + ((s0 + qi) * H + h) * D);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    lse[(int64_t)h * t_total + s0 + qi] = (m + log2f(l)) * kLn2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

inline bool attn_umma_supported(int dtype, int head_dim) {
  return head_dim == 16 && (dtype == B2PC_F16 || dtype == B2PC_BF16);
}

inline int attn_block_n() {
  static int bn = [] {
    const char* e = getenv("B2PC_ATTN_BN");
    int v = e ? atoi(e) : 64;
    return v == 128 ? 128 : 64;
  }();
  return bn;
}

inline AttnDesc attn_desc(int bn) {
  AttnDesc d{(uint32_t)kAuQ * 16, 128, (uint32_t)bn * 16, 128, 128, (uint32_t)bn * 16};
  if (const char* e = getenv("B2PC_ATTN_DESC")) {  // bring-up aid: "q_lbo,q_sbo,k_lbo,k_sbo,v_lbo,v_sbo"
    unsigned v[6];
    if (sscanf(e, "%u,%u,%u,%u,%u,%u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6)
      d = AttnDesc{v[0], v[1], v[2], v[3], v[4], v[5]};
  }
  return d;
}

template <typename T, int BN>
inline int launch_attn_fwd_umma_t(const void* qkv, const int32_t* cu, int n_seq, int max_seqlen, int64_t t, int H, float scale,
                                  void* out, float* lse, cudaStream_t stream) {
  static bool configured = false;
  constexpr int smem = attn_fwd_smem_bytes<BN>();
  if (!configured) {
    cudaFuncSetAttribute(attn_fwd_umma_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(max_seqlen, kAuQ), n_seq, H);
  attn_fwd_umma_kernel<T, BN><<<grid, kAuQ, smem, stream>>>((const T*)qkv, cu, t, H, scale, (T*)out, lse, attn_desc(BN));
  count_launches(1);
  B2PC_CHECK_LAUNCH("patch_attn_fwd(tcgen05)");
  return B2PC_OK;
}

inline int launch_attn_fwd_umma(const void* qkv, int dtype, const int32_t* cu, int n_seq, int max_seqlen, int64_t t, int H, int D,
                                float scale, void* out, float* lse, cudaStream_t stream) {
  (void)D;
  if (n_seq == 0 || t == 0) return B2PC_OK;
  const int bn = attn_block_n();
  if (dtype == B2PC_BF16)
    return bn == 64 ? launch_attn_fwd_umma_t<__nv_bfloat16, 64>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, stream)
                    : launch_attn_fwd_umma_t<__nv_bfloat16, 128>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, stream);
  return bn == 64 ? launch_attn_fwd_umma_t<__half, 64>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, stream)
                  : launch_attn_fwd_umma_t<__half, 128>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, stream);
}

// backward: SIMT kernels until the tcgen05 backward lands
inline size_t attn_bwd_umma_workspace_bytes(int64_t t, int H, int D) { return attn_bwd_workspace_bytes(t, H, D); }
inline int launch_attn_bwd_umma(const void* dout, const void* qkv, const void* out, const float* lse, int dtype, const int32_t* cu,
                                int n_seq, int max_seqlen, int64_t t, int H, int D, float scale, void* dqkv, void* ws,
                                cudaStream_t stream) {
  if (dtype == B2PC_F16)
    return launch_attn_bwd_simt<__half>(dout, qkv, out, lse, cu, n_seq, max_seqlen, t, H, D, scale, dqkv, ws, stream);
  return launch_attn_bwd_simt<__nv_bfloat16>(dout, qkv, out, lse, cu, n_seq, max_seqlen, t, H, D, scale, dqkv, ws, stream);
}

}  // namespace b2pc
