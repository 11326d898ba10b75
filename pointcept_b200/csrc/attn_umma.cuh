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
template <int BN> __host__ __device__ constexpr int attn_fwd_smem_bytes() { return kAuQ * 32 + kAuStages * BN * 64 + 128; }

// TMA = true: Q / K / V tiles arrive by cp.async.bulk.tensor (one elected thread, byte-counted mbarriers); false: cp.async.
// GATHER = true (serialized attention, ptv3m1:188,216 fused in): qkv holds POINT rows [N, 3, H, 16]; slot t of the padded patch
// sequence reads point row gidx[t] (= order[pad][t]) and writes its output to point row sidx[t] when sidx[t] >= 0 (the slot is
// the point's primary slot; borrowed filler slots have sidx < 0 and are dropped) -- the [order] gather and the [inverse]
// gather of the reference happen inside the tile loads / the epilogue and the padded qkv / out tensors never exist.
template <typename T, int BN, bool TMA, bool GATHER>
__global__ void __launch_bounds__(kAuQ)
attn_fwd_umma_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ cu, int64_t t_total, int H, float scale,
                     T* __restrict__ out, float* __restrict__ lse, AttnDesc dd, const __grid_constant__ CUtensorMap tmap_q,
                     const __grid_constant__ CUtensorMap tmap_kv, const int32_t* __restrict__ gidx, const int32_t* __restrict__ sidx) {
  static_assert(!(TMA && GATHER), "gathered rows are fetched with cp.async");
  using namespace umma;
  constexpr int D = 16;
  constexpr int TMEM_COLS = attn_tmem_cols<BN>();
  constexpr uint32_t COL_S = 0, COL_P = BN, COL_O = BN + BN / 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* q_s = smem;
  uint8_t* kv_s = smem + kAuQ * 32;
  uint64_t* bar = reinterpret_cast<uint64_t*>(kv_s + kAuStages * BN * 64);
  uint64_t* full = bar + 1;             // [kAuStages] K/V stage landed (TMA path)
  uint64_t* qbar = full + kAuStages;    // Q tile landed (TMA path)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int seq = blockIdx.y, h = blockIdx.z;
  const int64_t s0 = cu[seq];
  const int len = (int)(cu[seq + 1] - s0);
  const int q0 = blockIdx.x * kAuQ;
  if (q0 >= len) return;
  const int nblk = (len + BN - 1) / BN;
  const int64_t row_stride = (int64_t)3 * H * D;  // elements between consecutive tokens
  // Bulk tile loads fetch whole BN-row boxes: past the end of a ragged sequence they would bring in rows of the NEXT sequence,
  // and although their probabilities are zeroed, 0 x Inf/NaN in the PV product is not -- so ragged sequences take the
  // zero-filling cp.async path (per CTA decision; full patches, i.e. nearly all of them, take TMA).
  const bool tma = TMA && (len % BN == 0);

  if (warp == 0) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  if (tid == 0) {
    mbar_init(bar, 1);
    for (int s = 0; s < kAuStages; ++s) mbar_init(&full[s], 1);
    mbar_init(qbar, 1);
    fence_mbar_init();
  }
  auto tma_kv = [&](int blk, int stage) {
    uint8_t* ks = kv_s + stage * BN * 64;
    const int row = (int)(s0 + blk * BN);
    mbar_expect_tx(&full[stage], BN * 64);
    tma_load_2d(smem_u32(ks), &tmap_kv, (1 * H + h) * D, row, &full[stage]);
    tma_load_2d(smem_u32(ks + BN * 16), &tmap_kv, (1 * H + h) * D + 8, row, &full[stage]);
    tma_load_2d(smem_u32(ks + BN * 32), &tmap_kv, (2 * H + h) * D, row, &full[stage]);
    tma_load_2d(smem_u32(ks + BN * 48), &tmap_kv, (2 * H + h) * D + 8, row, &full[stage]);
  };

  // ---- loads: 16-byte pieces into "plane" layout: piece (row r, chunk c) -> c * rows*16 + r*16 ----------------------
  const int64_t row_base = GATHER ? 0 : s0;   // GATHER: rows are addressed through gidx
  const T* base_q = qkv + (row_base * 3 + 0) * H * D + h * D;
  const T* base_k = qkv + (row_base * 3 + 1) * H * D + h * D;
  const T* base_v = qkv + (row_base * 3 + 2) * H * D + h * D;
  const int32_t* gix = GATHER ? gidx + s0 : nullptr;
  if (tma) {
    if (tid == 0) {
      mbar_expect_tx(qbar, kAuQ * 32);
      tma_load_2d(smem_u32(q_s), &tmap_q, h * D, (int)(s0 + q0), qbar);
      tma_load_2d(smem_u32(q_s + kAuQ * 16), &tmap_q, h * D + 8, (int)(s0 + q0), qbar);
      tma_kv(0, 0);
      if (nblk > 1) tma_kv(1, 1);
    }
  } else {
    const int r = tid;
    const bool ok = q0 + r < len;
    const int64_t prow = GATHER ? (ok ? (int64_t)__ldg(gix + q0 + r) : 0) : (int64_t)(q0 + r);
    const T* src = base_q + prow * row_stride;
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
      const int64_t prow = GATHER ? (ok ? (int64_t)__ldg(gix + k0 + r) : 0) : (int64_t)(k0 + r);
      const T* src = (which ? base_v : base_k) + prow * row_stride + c * 8;
      cp_async16(smem_u32((which ? vs : ks) + c * BN * 16 + r * 16), ok ? src : base_k, ok);
    }
  };
  if (!tma) {
    load_kv(0, 0);
    cp_async_commit();
    if (nblk > 1) load_kv(1, 1);
    cp_async_commit();
  }

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
    if (!tma) {
      cp_async_wait<1>();
      fence_proxy_async();
      __syncthreads();
    }
    if (tid == 0) {
      if (tma) {
        if (j == 0) mbar_wait(qbar, 0);
        mbar_wait(&full[stage], (j / kAuStages) & 1);
      }
      tc_fence_after();
      mma_ss(tmem_base + COL_S, desc_q, make_smem_desc(smem_u32(ks), dd.k_lbo, dd.k_sbo), idesc_s, 0);
      mma_commit(bar);
    }
    mbar_wait(bar, j & 1);
    tc_fence_after();
    // S_j is complete, hence so is PV_{j-1}: its K/V stage is free again -> prefetch block j+2 into it
    if (tma) {
      if (tid == 0 && j + 2 < nblk) tma_kv(j + 2, (j + 2) % kAuStages);
    } else {
      if (j + 2 < nblk) load_kv(j + 2, (j + 2) % kAuStages);
      cp_async_commit();
    }
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
    // S row of this block -> registers once (BN <= 64) or per 32-column chunk twice (BN = 128: max pass, then exp pass)
    constexpr bool kKeep = BN <= 64;
    constexpr int NCH = BN / 32;
    uint32_t sr[kKeep ? NCH : 1][32];
    float mx = -INFINITY;
    if (kKeep) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) tmem_ld32(lane_base + COL_S + ch * 32, sr[ch]);
      tmem_ld_wait();
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      uint32_t (&r)[32] = sr[kKeep ? ch : 0];
      if (!kKeep) { tmem_ld32(lane_base + COL_S + ch * 32, r); tmem_ld_wait(); }
      if (!tail) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
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
    // P = exp2(c*S - m_new), row sum, pack to the MMA operand type, store to TMEM
    uint64_t lsum2 = pack_f2(0.f, 0.f);
    const uint64_t c2 = pack_f2(c, c), nm2 = pack_f2(-m_new, -m_new);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      uint32_t (&r)[32] = sr[kKeep ? ch : 0];
      if (!kKeep) { tmem_ld32(lane_base + COL_S + ch * 32, r); tmem_ld_wait(); }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float t0, t1;
        unpack_f2(ffma2(pack_f2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nm2), t0, t1);
        float p0 = ex2(t0), p1 = ex2(t1);
        if (tail) {
          if (ch * 32 + 2 * i >= nvalid) p0 = 0.f;
          if (ch * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
        }
        lsum2 = fadd2(lsum2, pack_f2(p0, p1));
        pk[i] = pack2<T>(p0, p1);
      }
      tmem_st16(lane_base + COL_P + ch * 16, pk);
    }
    {
      float la, lb;
      unpack_f2(lsum2, la, lb);
      l += la + lb;
    }
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
    const int64_t orow = GATHER ? (int64_t)__ldg(sidx + s0 + qi) : s0 + qi;
    if (orow >= 0) {
      uint4* dst = reinterpret_cast<uint4*>(out + (orow * H + h) * D);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
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

inline bool attn_use_tma() {
  static bool v = [] { const char* e = getenv("B2PC_ATTN_TMA"); return !(e && atoi(e) == 0); }();
  return v;
}

// gidx / sidx non-null: serialized (gather-fused) mode, see the kernel comment
template <typename T, int BN>
inline int launch_attn_fwd_umma_t(const void* qkv, const int32_t* cu, int n_seq, int max_seqlen, int64_t t, int H, float scale,
                                  void* out, float* lse, const int32_t* gidx, const int32_t* sidx, cudaStream_t stream) {
  constexpr int smem = attn_fwd_smem_bytes<BN>();
  dim3 grid((unsigned)ceil_div(max_seqlen, kAuQ), n_seq, H);
  CUtensorMap mq, mkv;
  if (gidx) {
    cudaFuncSetAttribute(attn_fwd_umma_kernel<T, BN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attn_fwd_umma_kernel<T, BN, false, true><<<grid, kAuQ, smem, stream>>>((const T*)qkv, cu, t, H, scale, (T*)out, lse, attn_desc(BN), mq, mkv,
                                                                           gidx, sidx);
  } else {
    const bool bf16 = UmmaFmt<T>::v == umma::kFmtBF16;
    const bool tma = attn_use_tma() && t < (1ll << 31) && make_plane_tensor_map(&mq, qkv, bf16, (uint64_t)t, (uint64_t)3 * H * 16, kAuQ) &&
                     make_plane_tensor_map(&mkv, qkv, bf16, (uint64_t)t, (uint64_t)3 * H * 16, BN);
    // the attribute is per device: set it on every launch (a few hundred ns) instead of caching it in a process-wide static
    if (tma) {
      cudaFuncSetAttribute(attn_fwd_umma_kernel<T, BN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attn_fwd_umma_kernel<T, BN, true, false><<<grid, kAuQ, smem, stream>>>((const T*)qkv, cu, t, H, scale, (T*)out, lse, attn_desc(BN), mq, mkv,
                                                                             nullptr, nullptr);
    } else {
      cudaFuncSetAttribute(attn_fwd_umma_kernel<T, BN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attn_fwd_umma_kernel<T, BN, false, false><<<grid, kAuQ, smem, stream>>>((const T*)qkv, cu, t, H, scale, (T*)out, lse, attn_desc(BN), mq,
                                                                              mkv, nullptr, nullptr);
    }
  }
  count_launches(1);
  B2PC_CHECK_LAUNCH("patch_attn_fwd(tcgen05)");
  return B2PC_OK;
}

inline int launch_attn_fwd_umma(const void* qkv, int dtype, const int32_t* cu, int n_seq, int max_seqlen, int64_t t, int H, int D,
                                float scale, void* out, float* lse, cudaStream_t stream, const int32_t* gidx = nullptr,
                                const int32_t* sidx = nullptr) {
  (void)D;
  if (n_seq == 0 || t == 0) return B2PC_OK;
  const int bn = attn_block_n();
  if (dtype == B2PC_BF16)
    return bn == 64 ? launch_attn_fwd_umma_t<__nv_bfloat16, 64>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, gidx, sidx, stream)
                    : launch_attn_fwd_umma_t<__nv_bfloat16, 128>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, gidx, sidx, stream);
  return bn == 64 ? launch_attn_fwd_umma_t<__half, 64>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, gidx, sidx, stream)
                  : launch_attn_fwd_umma_t<__half, 128>(qkv, cu, n_seq, max_seqlen, t, H, scale, out, lse, gidx, sidx, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward.  One CTA = one (sequence, head, block of 128 keys); thread t owns key row t (= TMEM lane t) and sweeps the
// queries in blocks of BQ = 64 columns, everything in the transposed frame so that the CTA-owned dK/dV accumulate in TMEM
// across the sweep and P / dS never leave the SM:
//     S^T  = K_j Q_i^T     M=128 N=64 K=16   (A = K_j smem, B = Q_i smem)                 -> TMEM [0,64)
//     dP^T = V_j dO_i^T    M=128 N=64 K=16   (A = V_j smem, B = dO_i smem)                -> TMEM [64,128)
//     P^T  = exp2(c S^T - lse2_i),  dS^T = P^T * (dP^T - delta_i)      (registers; bf16/fp16 copies to TMEM + smem)
//     dV_j += P^T  dO_i    M=128 N=16 K=64   (A = P^T  TMEM [128,160), B = dO_i tile read MN-major)  -> TMEM [192,208)
//     dK_j += dS^T Q_i     M=128 N=16 K=64   (A = dS^T TMEM [160,192), B = Q_i  tile read MN-major)  -> TMEM [208,224)
//     dQ_i  = dS   K_j     M=64  N=16 K=128  (A = dS smem MN-major,    B = K_j tile read MN-major)   -> TMEM [224,240)
// dQ_i partials are added into an fp32 accumulator with vector reductions (red.global.add.v4.f32) and converted at the end.
// The same shared-memory bytes serve as K-major operand (rows x 16 channels) and as MN-major operand (16 channels x rows):
// only the descriptor differs.
constexpr int kAbK = 128;   // keys per CTA (= threads)
// Q / dO ring depth ST.  ST = 3 (round-1 order): the loads of block i+2 are issued after the wait for block i's S / dP MMAs,
// because their stage was read by block i-1's dV / dK MMAs, which only that wait covers.  ST = 4: the stage of block i+2 was last
// read by block i-2, complete since the previous wait, so the loads (address arithmetic, and in serialized mode the dependent
// gidx / sidx reads, which are fetched one block further ahead into a register) are issued in the shadow of the MMA wait instead of
// on the block's critical path.
// BQ = queries per sweep step.  BQ = 64: 256 TMEM columns (S 64 | dP 64 | P 32 | dS 32 | dV dK dQ 48) -> 2 CTAs per SM.
// BQ = 32: P overwrites the S columns and dS the dP columns once a thread holds its row of both in registers (TMEM lanes are
// private to their thread), so 128 columns suffice (S/P 32 | dP/dS 32 | dV dK dQ 48) and FOUR CTAs share an SM: the kernel is
// bound by the latency of its issue -> commit -> wait -> softmax -> barrier chain (ncu: barrier + wait stalls dominate, XU pipe
// 27 %), which twice as many resident CTAs hide.  The dQ MMA keeps M = 64 (upper 32 rows of its operand are zero planes).
// smem: K_j 4096 | V_j 4096 | dS (A of the dQ MMA) 64x128x2 = 16384 | stages x (Q BQ*32 | dO BQ*32 | lse2 BQ*4 | delta BQ*4) | bar, slot
template <int BQ> __host__ __device__ constexpr int attn_bwd_stage_bytes() { return BQ * 32 + BQ * 32 + BQ * 4 + BQ * 4; }
template <int BQ, int ST> __host__ __device__ constexpr int attn_bwd_smem_bytes() { return 4096 + 4096 + 16384 + (ST == 3 ? 3 : 4) * attn_bwd_stage_bytes<BQ>() + 64; }
template <int BQ> __host__ __device__ constexpr int attn_bwd_tmem_cols() { return BQ == 64 ? 256 : 128; }

// GATHER = true: serialized mode (see the forward kernel): qkv / dout / dqkv hold POINT rows; slot t reads point row gidx[t],
// its dO is dout[sidx[t]] when sidx[t] >= 0 and zero otherwise (the output of a borrowed filler slot was dropped); dK / dV of a
// primary slot go straight to the point's row of dqkv, those of filler slot with sidx = -(r+1) to row r of `side` [n_dup, 2, H, 16]
// (added to the point's row afterwards: a point owns at most one filler slot besides its primary one).
template <typename T, bool GATHER, int BQ, int ST>
__global__ void __launch_bounds__(kAbK, BQ == 32 ? 4 : 2)
attn_bwd_umma_kernel(const T* __restrict__ dout, const T* __restrict__ qkv, const float* __restrict__ lse,
                     const float* __restrict__ delta, const int32_t* __restrict__ cu, int64_t t_total, int H, float scale,
                     T* __restrict__ dqkv, float* __restrict__ dq_acc, const int32_t* __restrict__ gidx,
                     const int32_t* __restrict__ sidx, T* __restrict__ side) {
  using namespace umma;
  constexpr int D = 16;
  static_assert(BQ == 32 || BQ == 64, "query block");
  static_assert(ST >= 3 && ST <= 6, "3 / 4: ring depth; 5: depth 4 + software-pipelined MMA order; 6: depth 4 + paired dQ MMA");
  static_assert(ST < 5 || BQ == 32, "modes 5 and 6 are written for one 32-query chunk per block");
  constexpr bool PIPE = ST == 5;
  // PAIR: the dQ MMA has M = 64 but a block brings 32 queries, so half of every dQ MMA multiplies zero planes.  Two consecutive
  // blocks write their dS into the lower / upper four planes of the operand and the 8 dQ MMAs are issued once per pair: 10 instead
  // of 14 tcgen05.mma per block (ncu: the TC pipe is busy 55 % of the time at ~36 cycles per instruction, whatever its N).
  constexpr bool PAIR = ST == 6;
  constexpr bool EARLY = ST == 4 || ST == 6;   // loads of block i+2 in the shadow of the MMA wait
  constexpr int kAbStages = ST == 3 ? 3 : 4;
  constexpr int kAbQ = BQ, kAbStageBytes = attn_bwd_stage_bytes<BQ>(), kAbTmemCols = attn_bwd_tmem_cols<BQ>();
  constexpr uint32_t COL_S = 0, COL_DP = BQ, COL_P = BQ == 64 ? 128 : 0, COL_DS = BQ == 64 ? 160 : BQ,
                     COL_DV = BQ == 64 ? 192 : 64, COL_DK = COL_DV + 16, COL_DQ = COL_DV + 32;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* k_s = smem;
  uint8_t* v_s = smem + 4096;
  uint8_t* ds_s = smem + 8192;
  uint8_t* st_s = smem + 8192 + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(st_s + kAbStages * kAbStageBytes);
  uint64_t* bar_b = bar + 1;                                   // pipelined order only: dQ MMAs of a block complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seq = blockIdx.y, h = blockIdx.z;
  const int64_t s0 = cu[seq];
  const int len = (int)(cu[seq + 1] - s0);
  const int k0 = blockIdx.x * kAbK;
  if (k0 >= len) return;
  const int nblk = (len + kAbQ - 1) / kAbQ;
  const int64_t row_stride = (int64_t)3 * H * D;

  if (warp == 0) { tmem_alloc(tmem_slot, kAbTmemCols); tmem_relinquish(); }
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar_b, 1); fence_mbar_init(); }

  const int64_t row_base = GATHER ? 0 : s0;
  const T* base_q = qkv + (row_base * 3 + 0) * H * D + h * D;
  const T* base_k = qkv + (row_base * 3 + 1) * H * D + h * D;
  const T* base_v = qkv + (row_base * 3 + 2) * H * D + h * D;
  const T* base_do = dout + row_base * H * D + h * D;
  const int32_t* gix = GATHER ? gidx + s0 : nullptr;
  const int32_t* six = GATHER ? sidx + s0 : nullptr;
  const float* base_lse = lse + (int64_t)h * t_total + s0;
  const float* base_dl = delta + (int64_t)h * t_total + s0;
  {  // K_j, V_j: thread = key row, two 16-byte pieces each, plane layout (chunk c -> c*2048 + row*16)
    const bool ok = k0 + tid < len;
    const int64_t prow = GATHER ? (ok ? (int64_t)__ldg(gix + k0 + tid) : 0) : (int64_t)(k0 + tid);
    const T* ks = base_k + prow * row_stride;
    const T* vs = base_v + prow * row_stride;
    cp_async16(smem_u32(k_s + tid * 16), ok ? ks : base_k, ok);
    cp_async16(smem_u32(k_s + 2048 + tid * 16), ok ? ks + 8 : base_k, ok);
    cp_async16(smem_u32(v_s + tid * 16), ok ? vs : base_k, ok);
    cp_async16(smem_u32(v_s + 2048 + tid * 16), ok ? vs + 8 : base_k, ok);
  }
  // row index a loading thread needs for query block blk: gidx (Q rows) for threads 0-63, sidx (dO rows, < 0 = filler slot) for
  // threads 64-127; packed mode: the slot itself
  auto fetch_idx = [&](int blk) -> int64_t {
    const int which = tid >> 6, q = blk * kAbQ + (tid & 63);
    const bool ok = (tid & 63) < kAbQ && blk < nblk && q < len;
    if (which == 0) return ok ? (GATHER ? (int64_t)__ldg(gix + q) : (int64_t)q) : 0;
    return ok ? (GATHER ? (int64_t)__ldg(six + q) : (int64_t)q) : -1;
  };
  auto load_q = [&](int blk, int stage, int64_t prow) {
    uint8_t* st = st_s + stage * kAbStageBytes;
    const int q0 = blk * kAbQ;
    if ((tid & 63) < kAbQ) {  // BQ rows x 2 chunks of Q and of dO: thread -> (which, row)
      const int which = tid >> 6, r = (tid & 63);
      const bool ok = q0 + r < len;
      if (which == 0) {
        const T* src = base_q + prow * row_stride;
        cp_async16(smem_u32(st + r * 16), ok ? src : base_q, ok);
        cp_async16(smem_u32(st + kAbQ * 16 + r * 16), ok ? src + 8 : base_q, ok);
      } else {
        const bool okd = ok && prow >= 0;
        const T* src = base_do + (okd ? prow : 0) * (H * D);
        cp_async16(smem_u32(st + kAbQ * 32 + r * 16), src, okd);
        cp_async16(smem_u32(st + kAbQ * 32 + kAbQ * 16 + r * 16), src + 8, okd);
      }
      // lse / delta of the BQ queries (4-byte async copies; zero when the query does not exist)
      const float* g = (which == 0 ? base_lse : base_dl) + q0 + r;
      cp_async4(smem_u32(st + kAbQ * 64 + which * (kAbQ * 4) + r * 4), ok ? g : base_lse, ok);
    }
  };
  load_q(0, 0, fetch_idx(0));
  cp_async_commit();
  if (nblk > 1) load_q(1, 1, fetch_idx(1));
  cp_async_commit();
  int64_t idx_next = ST >= 4 ? fetch_idx(2) : 0;   // modes 4-6   // ST = 4: index of the block loaded during the next sweep step
  if (kAbQ < 64) {   // the dQ MMA has M = 64: the query rows this kernel never fills are zero planes of its A operand
    for (int q = tid; q < (64 - kAbQ) * 128 * 2 / 16; q += kAbK) reinterpret_cast<uint4*>(ds_s + kAbQ * 256)[q] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);

  constexpr uint32_t idesc_s = make_idesc(128, kAbQ, UmmaFmt<T>::v, UmmaFmt<T>::v, 0, 0);   // S^T, dP^T
  constexpr uint32_t idesc_kv = make_idesc(128, 16, UmmaFmt<T>::v, UmmaFmt<T>::v, 0, 1);    // dV, dK: B MN-major
  constexpr uint32_t idesc_dq = make_idesc(64, 16, UmmaFmt<T>::v, UmmaFmt<T>::v, 1, 1);     // dQ: A and B MN-major
  const uint64_t desc_k = make_smem_desc(smem_u32(k_s), 2048, 128);
  const uint64_t desc_v = make_smem_desc(smem_u32(v_s), 2048, 128);
  const float c = scale * kLog2e;
  const uint64_t c2 = pack_f2(c, c);

  auto flush_dq = [&](int blk) {
    // dQ partial of query block blk (PAIR: of the two blocks 2*blk, 2*blk+1): M=64 accumulator, row r lives in lane
    // (r%16) + 32*(r/16); 16 fp32 columns
    constexpr int kRows = PAIR ? 64 : kAbQ;
    uint32_t r[16];
    tmem_ld16(lane_base + COL_DQ, r);
    tmem_ld_wait();
    const int qi = blk * kRows + warp * 16 + lane;
    if (lane < 16 && warp * 16 < kRows && qi < len) {
      float* dst = dq_acc + ((s0 + qi) * H + h) * D;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        red_add_v4(dst + 4 * i, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                   __uint_as_float(r[4 * i + 3]));
    }
  };

  // P^T, dS^T of one query block from S^T / dP^T in TMEM: written back to TMEM (A operands of the dV / dK MMAs) and, dS, to
  // shared memory (A operand of the dQ MMA); pre_store() runs once before the first shared-memory store
  auto softmax_block = [&](uint8_t* st, int plane0, auto&& pre_store) {
    const float* lse_s = reinterpret_cast<const float*>(st + kAbQ * 64);
    const float* dl_s = reinterpret_cast<const float*>(st + kAbQ * 64 + kAbQ * 4);
#pragma unroll
    for (int ch = 0; ch < kAbQ / 32; ++ch) {
      uint32_t s_r[32], dp_r[32];
      tmem_ld32(lane_base + COL_S + ch * 32, s_r);
      tmem_ld32(lane_base + COL_DP + ch * 32, dp_r);
      tmem_ld_wait();
      uint32_t pp[16], dd[16];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        // lse_s holds -lse*log2(e), dl_s holds -delta (prepared by attn_delta_kernel): two queries per packed instruction
        const ulonglong2 l4 = *reinterpret_cast<const ulonglong2*>(lse_s + ch * 32 + g * 4);
        const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(dl_s + ch * 32 + g * 4);
        float p[4], ds[4];
        {
          const uint64_t t = ffma2(pack_f2(__uint_as_float(s_r[g * 4]), __uint_as_float(s_r[g * 4 + 1])), c2, l4.x);
          float t0, t1;
          unpack_f2(t, t0, t1);
          p[0] = ex2(t0); p[1] = ex2(t1);
          unpack_f2(fmul2(pack_f2(p[0], p[1]), fadd2(pack_f2(__uint_as_float(dp_r[g * 4]), __uint_as_float(dp_r[g * 4 + 1])), d4.x)),
                    ds[0], ds[1]);
        }
        {
          const uint64_t t = ffma2(pack_f2(__uint_as_float(s_r[g * 4 + 2]), __uint_as_float(s_r[g * 4 + 3])), c2, l4.y);
          float t0, t1;
          unpack_f2(t, t0, t1);
          p[2] = ex2(t0); p[3] = ex2(t1);
          unpack_f2(fmul2(pack_f2(p[2], p[3]), fadd2(pack_f2(__uint_as_float(dp_r[g * 4 + 2]), __uint_as_float(dp_r[g * 4 + 3])), d4.y)),
                    ds[2], ds[3]);
        }
        pp[g * 2] = pack2<T>(p[0], p[1]);
        pp[g * 2 + 1] = pack2<T>(p[2], p[3]);
        dd[g * 2] = pack2<T>(ds[0], ds[1]);
        dd[g * 2 + 1] = pack2<T>(ds[2], ds[3]);
      }
      tmem_st16(lane_base + COL_P + ch * 16, pp);
      tmem_st16(lane_base + COL_DS + ch * 16, dd);
      if (ch == 0) pre_store();   // pipelined order: the previous block's dQ MMA must be done reading ds_s
      // dS as the (MN-major) A operand of the dQ MMA: 8-query piece p of this key -> p*2048 + key*16
#pragma unroll
      for (int pc = 0; pc < 4; ++pc)
        *reinterpret_cast<uint4*>(ds_s + (plane0 + ch * 4 + pc) * 2048 + tid * 16) =
            make_uint4(dd[pc * 4], dd[pc * 4 + 1], dd[pc * 4 + 2], dd[pc * 4 + 3]);
    }
  };
  auto issue_sdp = [&](uint8_t* st) {      // S^T = K Q^T and dP^T = V dO^T of the query block staged at st (thread 0 only)
    mma_ss(tmem_base + COL_S, desc_k, make_smem_desc(smem_u32(st), kAbQ * 16, 128), idesc_s, 0);
    mma_ss(tmem_base + COL_DP, desc_v, make_smem_desc(smem_u32(st + kAbQ * 32), kAbQ * 16, 128), idesc_s, 0);
  };
  auto issue_dvdk = [&](uint8_t* st, bool first) {   // reduction over the BQ queries, 16 per MMA
#pragma unroll
    for (int ks = 0; ks < kAbQ / 16; ++ks) {
      mma_ts(tmem_base + COL_DV, tmem_base + COL_P + ks * 8, make_smem_desc(smem_u32(st + kAbQ * 32 + ks * 256), 128, kAbQ * 16),
             idesc_kv, (!first || ks > 0) ? 1u : 0u);
      mma_ts(tmem_base + COL_DK, tmem_base + COL_DS + ks * 8, make_smem_desc(smem_u32(st + ks * 256), 128, kAbQ * 16), idesc_kv,
             (!first || ks > 0) ? 1u : 0u);
    }
  };
  auto issue_dq = [&]() {                   // reduction over the 128 keys
#pragma unroll
    for (int ks = 0; ks < kAbK / 16; ++ks)
      mma_ss(tmem_base + COL_DQ, make_smem_desc(smem_u32(ds_s + ks * 256), 128, 2048),
             make_smem_desc(smem_u32(k_s + ks * 256), 128, 2048), idesc_dq, ks > 0 ? 1u : 0u);
  };

  if constexpr (PIPE) {
    // Software-pipelined order (one block barrier per query block instead of two): S / dP of block i+1 are issued right behind the
    // dV / dK MMAs of block i, and the dQ MMA behind their commit on a second mbarrier, so the wait at the top of a block covers
    // only what the softmax needs and the 8 dQ MMAs run under the TMEM loads and the exponentials of the next block.
    //   barA phase k: S / dP of block k (and dV / dK of block k-1) complete;  barB phase k: dQ of block k complete.
    cp_async_wait<1>();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_sdp(st_s);
      mma_commit(bar);
    }
    for (int i = 0; i < nblk; ++i) {
      uint8_t* st = st_s + (i % kAbStages) * kAbStageBytes;
      if (i + 2 < nblk) load_q(i + 2, (i + 2) % kAbStages, idx_next);   // its stage was released by the wait of block i-1
      cp_async_commit();
      idx_next = fetch_idx(i + 3);
      mbar_wait(bar, i & 1);
      tc_fence_after();
      softmax_block(st, 0, [&] {
        if (i > 0) {
          mbar_wait(bar_b, (i - 1) & 1);
          tc_fence_after();
          flush_dq(i - 1);
        }
      });
      tmem_st_wait();
      cp_async_wait<1>();          // block i+1 has landed (its S / dP MMAs are issued below)
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        issue_dvdk(st, i == 0);
        if (i + 1 < nblk) issue_sdp(st_s + ((i + 1) % kAbStages) * kAbStageBytes);
        mma_commit(bar);
        issue_dq();
        mma_commit(bar_b);
      }
    }
    mbar_wait(bar, nblk & 1);
    mbar_wait(bar_b, (nblk - 1) & 1);
    tc_fence_after();
    flush_dq(nblk - 1);
  } else {
  for (int i = 0; i < nblk; ++i) {
    const int stage = i % kAbStages;
    uint8_t* st = st_s + stage * kAbStageBytes;
    cp_async_wait<1>();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_sdp(st);
      mma_commit(bar);
    }
    if (EARLY) {   // stage (i+2)%4 was released by the previous wait: load in the shadow of this one
      if (i + 2 < nblk) load_q(i + 2, (i + 2) % kAbStages, idx_next);
      cp_async_commit();
      idx_next = fetch_idx(i + 3);
    }
    mbar_wait(bar, i & 1);
    tc_fence_after();
    if (ST == 3) {
      if (i + 2 < nblk) load_q(i + 2, (i + 2) % kAbStages, fetch_idx(i + 2));
      cp_async_commit();
    }
    if (PAIR) {
      if (i > 0 && ((i - 1) & 1)) flush_dq((i - 1) >> 1);     // the pair that ended with block i-1
    } else if (i > 0) {
      flush_dq(i - 1);
    }
    softmax_block(st, PAIR ? (i & 1) * 4 : 0, [] {});
    tmem_st_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_dvdk(st, i == 0);
      if (!PAIR || (i & 1) || i == nblk - 1) issue_dq();
      if (i == nblk - 1) mma_commit(bar);
    }
  }
  mbar_wait(bar, nblk & 1);
  tc_fence_after();
  flush_dq(PAIR ? (nblk - 1) >> 1 : nblk - 1);
  }
  {
    uint32_t rv[16], rk[16];
    tmem_ld16(lane_base + COL_DV, rv);
    tmem_ld16(lane_base + COL_DK, rk);
    tmem_ld_wait();
    const int ki = k0 + tid;
    if (ki < len) {
      uint32_t wv[8], wk[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        wv[e] = pack2<T>(__uint_as_float(rv[2 * e]), __uint_as_float(rv[2 * e + 1]));
        wk[e] = pack2<T>(__uint_as_float(rk[2 * e]) * scale, __uint_as_float(rk[2 * e + 1]) * scale);
      }
      uint4 *dk, *dv;
      if (GATHER) {
        const int64_t si = __ldg(six + ki);
        if (si >= 0) {
          dk = reinterpret_cast<uint4*>(dqkv + ((si * 3 + 1) * H + h) * D);
          dv = reinterpret_cast<uint4*>(dqkv + ((si * 3 + 2) * H + h) * D);
        } else {
          dk = reinterpret_cast<uint4*>(side + (((-si - 1) * 2 + 0) * H + h) * D);
          dv = reinterpret_cast<uint4*>(side + (((-si - 1) * 2 + 1) * H + h) * D);
        }
      } else {
        dk = reinterpret_cast<uint4*>(dqkv + (((s0 + ki) * 3 + 1) * H + h) * D);
        dv = reinterpret_cast<uint4*>(dqkv + (((s0 + ki) * 3 + 2) * H + h) * D);
      }
      dk[0] = make_uint4(wk[0], wk[1], wk[2], wk[3]);
      dk[1] = make_uint4(wk[4], wk[5], wk[6], wk[7]);
      dv[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      dv[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kAbTmemCols);
}

// dqkv[row(t), 0, h, :] = dq_acc[t, h, :] * scale; serialized mode: row(t) = sidx[t] (filler slots have a zero dQ and are skipped)
template <typename T>
__global__ void __launch_bounds__(256)
attn_dq_finish_kernel(const float* __restrict__ dq_acc, int64_t n_rows /* T*H */, int H, float scale, T* __restrict__ dqkv,
                      const int32_t* __restrict__ sidx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  int64_t t = i / H;
  const int h = (int)(i % H);
  if (sidx) {
    t = sidx[t];
    if (t < 0) return;
  }
  const float4* src = reinterpret_cast<const float4*>(dq_acc + i * 16);
  uint32_t w[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 v = src[e];
    w[2 * e] = pack2<T>(v.x * scale, v.y * scale);
    w[2 * e + 1] = pack2<T>(v.z * scale, v.w * scale);
  }
  uint4* dst = reinterpret_cast<uint4*>(dqkv + ((t * 3 + 0) * H + h) * 16);
  dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
  dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// serialized mode: dqkv[dup_point[r], 1 + which, h, :] += side[r, which, h, :]   (one thread per (r, which, h), 16 channels)
template <typename T>
__global__ void __launch_bounds__(256)
attn_dup_add_kernel(const T* __restrict__ side, const int32_t* __restrict__ dup_point, int64_t n_dup, int H, T* __restrict__ dqkv) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // over n_dup * 2 * H
  if (i >= n_dup * 2 * H) return;
  const int h = (int)(i % H);
  const int which = (int)((i / H) % 2);
  const int64_t r = i / (2 * H);
  const T* src = side + i * 16;
  T* dst = dqkv + (((int64_t)dup_point[r] * 3 + 1 + which) * H + h) * 16;
#pragma unroll
  for (int e = 0; e < 16; ++e) dst[e] = from_f32<T>(to_f32(dst[e]) + to_f32(src[e]));
}

inline size_t attn_bwd_umma_workspace_bytes(int64_t t, int H, int D, int64_t n_dup = 0) {
  (void)D;
  return 2 * align_up((size_t)t * H * sizeof(float), 256) + align_up((size_t)t * H * 16 * sizeof(float), 256) +
         align_up((size_t)n_dup * 2 * H * 16 * 2, 256) + 256;
}

// ndelta[h, t] = -sum_d dout*out,  nlse2[h, t] = -lse[h, t] * log2(e)   (the signs / scale the main kernel's packed FMAs want);
// serialized mode: dout / out are point rows, slot t uses row sidx[t] and gets delta = 0 when it is a filler slot
template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const T* __restrict__ dout, const T* __restrict__ out, const float* __restrict__ lse, int64_t t_total, int H,
                     float* __restrict__ ndelta, float* __restrict__ nlse2, const int32_t* __restrict__ sidx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // over T*H, (t, h) order
  if (i >= t_total * H) return;
  const int64_t t = i / H;
  const int h = (int)(i % H);
  float acc = 0.f;
  const int64_t row = sidx ? (int64_t)sidx[t] : t;
  if (row >= 0) {
    const uint4* a = reinterpret_cast<const uint4*>(dout + (row * H + h) * 16);
    const uint4* b = reinterpret_cast<const uint4*>(out + (row * H + h) * 16);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint4 va = a[q], vb = b[q];
      const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const T* pa = reinterpret_cast<const T*>(&wa[e]);
        const T* pb = reinterpret_cast<const T*>(&wb[e]);
        acc = fmaf(to_f32(pa[0]), to_f32(pb[0]), acc);
        acc = fmaf(to_f32(pa[1]), to_f32(pb[1]), acc);
      }
    }
  }
  ndelta[(int64_t)h * t_total + t] = -acc;
  nlse2[(int64_t)h * t_total + t] = -lse[(int64_t)h * t_total + t] * kLog2e;
}

// gidx non-null: serialized mode (dout / qkv / out / dqkv are point rows; see the kernel comments)
template <typename T>
inline int launch_attn_bwd_umma_t(const void* dout, const void* qkv, const void* out, const float* lse, const int32_t* cu, int n_seq,
                                  int max_seqlen, int64_t t, int H, float scale, void* dqkv, void* ws, cudaStream_t stream,
                                  const int32_t* gidx, const int32_t* sidx, const int32_t* dup_point, int64_t n_dup) {
  float* delta = (float*)ws;                                                        // holds -delta
  float* nlse2 = (float*)((char*)ws + align_up((size_t)t * H * sizeof(float), 256));  // holds -lse * log2(e)
  float* dq_acc = (float*)((char*)ws + 2 * align_up((size_t)t * H * sizeof(float), 256));
  T* side = (T*)((char*)dq_acc + align_up((size_t)t * H * 16 * sizeof(float), 256));
  attn_bwd_prep_kernel<T><<<(unsigned)ceil_div(t * H, 256), 256, 0, stream>>>((const T*)dout, (const T*)out, lse, t, H, delta, nlse2, sidx);
  cudaMemsetAsync(dq_acc, 0, (size_t)t * H * 16 * sizeof(float), stream);
  dim3 grid((unsigned)ceil_div(max_seqlen, kAbK), n_seq, H);
  static const int bq = [] { const char* e = getenv("B2PC_ATTN_BQ"); return (e && atoi(e) == 64) ? 64 : 32; }();
  // B2PC_ATTN_RING, measured on B200 at H = 2, T = 241 664, K = 1024 (profiles/r02_attn_bwd_ring_ab.txt):
  //   3: round-1 load order (0.441 ms); 4: loads in the shadow of the MMA wait (0.415 ms); 5: 4 + software-pipelined MMA order
  //   (0.433 ms: the block barrier was not the limiter); 6 (default): 4 + one dQ MMA burst per pair of query blocks (0.388 ms)
  static const int ring = [] { const char* e = getenv("B2PC_ATTN_RING"); const int v = e ? atoi(e) : 6; return (v == 3 || v == 4 || v == 5) ? v : 6; }();
#define B2PC_ATTN_BWD_LAUNCH(G, Q, S)                                                                                                     \
  do {                                                                                                                                   \
    cudaFuncSetAttribute(attn_bwd_umma_kernel<T, G, Q, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_bwd_smem_bytes<Q, S>());    \
    attn_bwd_umma_kernel<T, G, Q, S><<<grid, kAbK, attn_bwd_smem_bytes<Q, S>(), stream>>>((const T*)dout, (const T*)qkv, nlse2, delta, cu, \
                                                                                         t, H, scale, (T*)dqkv, dq_acc, gidx, sidx, side); \
  } while (0)
  if (bq == 64) {   // round-1 tiling (2 CTAs per SM), kept for A/B
    if (gidx) B2PC_ATTN_BWD_LAUNCH(true, 64, 3); else B2PC_ATTN_BWD_LAUNCH(false, 64, 3);
  } else if (ring == 3) {
    if (gidx) B2PC_ATTN_BWD_LAUNCH(true, 32, 3); else B2PC_ATTN_BWD_LAUNCH(false, 32, 3);
  } else if (ring == 5) {
    if (gidx) B2PC_ATTN_BWD_LAUNCH(true, 32, 5); else B2PC_ATTN_BWD_LAUNCH(false, 32, 5);
  } else if (ring == 6) {
    if (gidx) B2PC_ATTN_BWD_LAUNCH(true, 32, 6); else B2PC_ATTN_BWD_LAUNCH(false, 32, 6);
  } else {
    if (gidx) B2PC_ATTN_BWD_LAUNCH(true, 32, 4); else B2PC_ATTN_BWD_LAUNCH(false, 32, 4);
  }
#undef B2PC_ATTN_BWD_LAUNCH
  attn_dq_finish_kernel<T><<<(unsigned)ceil_div(t * H, 256), 256, 0, stream>>>(dq_acc, t * H, H, scale, (T*)dqkv, sidx);
  count_launches(3);
  if (gidx && n_dup > 0) {
    attn_dup_add_kernel<T><<<(unsigned)ceil_div(n_dup * 2 * H, 256), 256, 0, stream>>>(side, dup_point, n_dup, H, (T*)dqkv);
    count_launches(1);
  }
  B2PC_CHECK_LAUNCH("patch_attn_bwd(tcgen05)");
  return B2PC_OK;
}

inline int launch_attn_bwd_umma(const void* dout, const void* qkv, const void* out, const float* lse, int dtype, const int32_t* cu,
                                int n_seq, int max_seqlen, int64_t t, int H, int D, float scale, void* dqkv, void* ws,
                                cudaStream_t stream, const int32_t* gidx = nullptr, const int32_t* sidx = nullptr,
                                const int32_t* dup_point = nullptr, int64_t n_dup = 0) {
  (void)D;
  if (n_seq == 0 || t == 0) return B2PC_OK;
  if (dtype == B2PC_F16)
    return launch_attn_bwd_umma_t<__half>(dout, qkv, out, lse, cu, n_seq, max_seqlen, t, H, scale, dqkv, ws, stream, gidx, sidx, dup_point, n_dup);
  return launch_attn_bwd_umma_t<__nv_bfloat16>(dout, qkv, out, lse, cu, n_seq, max_seqlen, t, H, scale, dqkv, ws, stream, gidx, sidx, dup_point,
                                               n_dup);
}

}  // namespace b2pc
