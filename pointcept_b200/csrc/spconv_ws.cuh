// Sparse convolution on tcgen05, second generation: persistent, warp-specialised kernels (producer warps / one MMA-issuing
// thread / epilogue warps talking through mbarriers only -- no block-wide barrier inside the main loops).
//
//   conv_ws_kernel   forward and backward-data:  out[j, :] = bias + sum_k  feat[pair[k', j], :] @ W_k       (output stationary)
//   wgrad_ws_kernel  weight gradient:            dW[co, k, ci] = sum_j dout[j, co] * feat[pair[k, j], ci]    (weight stationary)
//
// Both keep their fp32 accumulators in TMEM for the whole reduction, gather rows with cp.async (zero fill for absent pairs)
// into un-swizzled "plane" operand tiles (umma.cuh) through a deep ring so that tens of KB of gathers are in flight per SM,
// and are deterministic (no atomics: the forward writes every output once, the weight gradient reduces row-range splits in
// a fixed order).
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "attn_umma.cuh"   // UmmaFmt, pack2

namespace b2pc {

// ---- small helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int n) {   // wait until at most n of this thread's groups are pending
  switch (n) {
    case 0: umma::cp_async_wait<0>(); break;
    case 1: umma::cp_async_wait<1>(); break;
    case 2: umma::cp_async_wait<2>(); break;
    case 3: umma::cp_async_wait<3>(); break;
    case 4: umma::cp_async_wait<4>(); break;
    case 5: umma::cp_async_wait<5>(); break;
    case 6: umma::cp_async_wait<6>(); break;
    case 7: umma::cp_async_wait<7>(); break;
    case 8: umma::cp_async_wait<8>(); break;
    case 9: umma::cp_async_wait<9>(); break;
    case 10: umma::cp_async_wait<10>(); break;
    case 11: umma::cp_async_wait<11>(); break;
    default: umma::cp_async_wait<12>(); break;
  }
}
__device__ __forceinline__ uint32_t ld_shared_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(umma::smem_u32(p)));
  return v;
}

constexpr int kWsRows = 128;        // output rows per tile (= TMEM lanes)
constexpr int kWsThreads = 288;     // warps 0-3 epilogue, warps 4-7 producers (thread = row), warp 8 MMA
constexpr int kWsMaxStages = 16;
constexpr int kWsCtrlBytes = 512;   // barriers + bookkeeping at the start of dynamic shared memory
constexpr int kWsSmemBudget = 200 * 1024;
constexpr int kWsMaxKV = 128;       // kernel volume limit of this path (125 = 5^3 stem)
constexpr uint32_t kMetaEnd = 1u << 21;

struct ConvWsCfg {
  int kc, n_cc, n_tile, n_ntiles, stages, lag, tmem_cols, w_resident, a_bytes, b_bytes, stage_bytes, res_bytes, idx_bytes, smem_bytes, grid;
  long long n_items;
};

inline bool conv_ws_supported(int dtype, int c_in, int c_out, int kv) {
  if (dtype != B2PC_F16 && dtype != B2PC_BF16) return false;
  if (c_in % 16 != 0 || c_out % 16 != 0) return false;
  if (kv > kWsMaxKV) return false;
  return true;
}

inline ConvWsCfg conv_ws_cfg(int64_t n_out, int c_in, int c_out, int kv) {
  ConvWsCfg c;
  c.kc = c_in % 64 == 0 ? 64 : (c_in % 32 == 0 ? 32 : 16);
  c.n_cc = c_in / c.kc;
  const int64_t m_tiles = ceil_div(n_out > 0 ? n_out : 1, kWsRows);
  // widest N tile (rows are gathered once per N tile) that still gives every SM a work item; narrow levels fall back to
  // smaller tiles instead of splitting the reduction, so every output element is produced by exactly one CTA
  int best = 0;
  for (int nt = c_out <= 256 ? c_out : 256; nt >= 16; nt -= 16) {
    if (c_out % nt != 0) continue;
    best = nt;
    if (m_tiles * (c_out / nt) >= kNumSMs || nt <= 32) break;
  }
  c.n_tile = best;
  c.n_ntiles = c_out / c.n_tile;
  c.n_items = m_tiles * c.n_ntiles;
  c.a_bytes = kWsRows * c.kc * 2;
  c.b_bytes = c.n_tile * c.kc * 2;
  c.idx_bytes = 2 * kv * kWsRows * 4 + 2 * kWsMaxKV;       // two rulebook-slice buffers + two active-offset lists
  const long long res = (long long)kv * c.n_cc * c.b_bytes;
  const long long avail = kWsSmemBudget - kWsCtrlBytes - c.idx_bytes;
  c.w_resident = (c.n_ntiles == 1 && res + 8 * c.a_bytes <= avail) ? 1 : 0;
  c.res_bytes = c.w_resident ? (int)res : 0;
  c.stage_bytes = c.a_bytes + (c.w_resident ? 0 : c.b_bytes);
  int st = (int)((avail - c.res_bytes) / c.stage_bytes);
  if (st > kWsMaxStages) st = kWsMaxStages;
  if (st < 3) st = 3;
  c.stages = st;
  c.lag = st >= 8 ? 6 : 2;
  int ncol = 32;                        // accumulator buffers sit on power-of-two column strides (N = 96 tiles at column 96 fault)
  while (ncol < c.n_tile) ncol <<= 1;
  c.tmem_cols = 2 * ncol;
  c.smem_bytes = kWsCtrlBytes + c.idx_bytes + c.res_bytes + c.stages * c.stage_bytes;
  c.grid = (int)(c.n_items < kNumSMs ? c.n_items : kNumSMs);
  return c;
}

struct WsCtrl {            // lives at the start of dynamic smem
  uint64_t full[kWsMaxStages];
  uint64_t empty[kWsMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t meta[kWsMaxStages];
  uint32_t mask_x[3][4];
  uint32_t tmem_slot;
};
static_assert(sizeof(WsCtrl) <= kWsCtrlBytes, "control block too large");

// Producer threads own one output row each (128 rows = 4 warps) and copy that row's KC channels of every active kernel offset with
// KC/8 cp.async from ONE base pointer -- the leanest instruction stream per gathered byte, which is what bounds this kernel
// (ncu: with one warp per scheduler the producers are issue-latency bound).  The rulebook slice of the NEXT tile is fetched by
// 4-byte cp.async into a second shared-memory buffer while the current tile's rows stream, so no index load is ever exposed.
template <typename T, int KC, int LAG /* unused: hand-over is wait-free */>
__global__ void __launch_bounds__(kWsThreads, 1)
conv_ws_kernel(const T* __restrict__ feat, const T* __restrict__ weight, const T* __restrict__ bias,
               const int32_t* __restrict__ pair, int64_t pair_stride, int64_t n_out, int c_in, int c_out, int kv,
               int transpose_w, int flip, T* __restrict__ out, ConvWsCfg cfg) {
  using namespace umma;
  extern __shared__ __align__(128) uint8_t smem[];
  WsCtrl* ctl = reinterpret_cast<WsCtrl*>(smem);
  int32_t* idx_s = reinterpret_cast<int32_t*>(smem + kWsCtrlBytes);                       // [2][kv][128]
  uint8_t* act_s = smem + kWsCtrlBytes + 2 * kv * kWsRows * 4;                            // [2][kWsMaxKV]
  uint8_t* w_res = smem + kWsCtrlBytes + cfg.idx_bytes;
  uint8_t* ring = w_res + cfg.res_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = cfg.stages, n_tile = cfg.n_tile, n_cc = cfg.n_cc, n_ntiles = cfg.n_ntiles;
  const int a_bytes = cfg.a_bytes, b_bytes = cfg.b_bytes, stage_bytes = cfg.stage_bytes;
  const bool resident = cfg.w_resident != 0;
  const long long n_items = cfg.n_items;
  constexpr int LPR = KC / 8;      // 16-byte pieces per gathered row

  if (warp == 0) { tmem_alloc(&ctl->tmem_slot, cfg.tmem_cols); tmem_relinquish(); }
  if (tid == 32) {
    for (int s = 0; s < S; ++s) { mbar_init(&ctl->full[s], 129); mbar_init(&ctl->empty[s], 1); }   // 128 copy-tracking arrivals + 1
    for (int b = 0; b < 2; ++b) { mbar_init(&ctl->acc_full[b], 1); mbar_init(&ctl->acc_empty[b], 128); }
    for (int i = 0; i < 12; ++i) (&ctl->mask_x[0][0])[i] = 0;
    fence_mbar_init();
  }
  // weight tile (kidx, cc) as a B operand: K-major [n_tile x KC] (forward) or MN-major [KC x n_tile] (backward data)
  auto load_w_tile = [&](uint8_t* dst, int kidx, int cc, int n0, int t0, int nthr) {
    if (!transpose_w) {
      for (int q = t0; q < n_tile * LPR; q += nthr) {
        const int n = q / LPR, p = q % LPR;
        cp_async16(smem_u32(dst) + p * (n_tile * 16) + n * 16, weight + ((int64_t)(n0 + n) * kv + kidx) * c_in + cc * KC + p * 8, true);
      }
    } else {
      const int ppr = n_tile / 8;
      for (int q = t0; q < KC * ppr; q += nthr) {
        const int kk = q / ppr, p = q % ppr;
        cp_async16(smem_u32(dst) + p * (KC * 16) + kk * 16, weight + ((int64_t)(cc * KC + kk) * kv + kidx) * c_out + n0 + p * 8, true);
      }
    }
  };
  if (resident) {   // all of W for this CTA's (only) N tile stays in shared memory for the whole kernel
    for (int wt = 0; wt < kv * n_cc; ++wt) load_w_tile(w_res + (size_t)wt * b_bytes, wt / n_cc, wt % n_cc, 0, tid, kWsThreads);
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_volatile_u32(&ctl->tmem_slot);

  if (warp >= 4 && warp < 8) {
    // ===================================== producers (thread = output row) =====================================
    const int ptid = tid - 128;
    // rulebook slice of one tile -> idx buffer `buf` (own row only; rows past n_out read a clamped address and are masked later)
    auto fetch_idx = [&](long long item, int buf) {
      int64_t j = (item / n_ntiles) * kWsRows + ptid;
      if (j >= n_out) j = n_out - 1;
      const uint32_t dst = smem_u32(idx_s + (size_t)buf * kv * kWsRows + ptid);
      for (int k = 0; k < kv; ++k) {
        const int kp = flip ? kv - 1 - k : k;
        cp_async4(dst + k * (kWsRows * 4), pair + (int64_t)kp * pair_stride + j, true);
      }
    };
    uint32_t gi = 0;                    // units (incl. END markers) produced so far
    int st_i = 0, st_ph = 0;            // ring position of the next unit: stage and ring-turn parity
    // A unit is handed over without any wait on the producer side: every thread lets the stage's `full` barrier track its own
    // outstanding copies (cp.async.mbarrier.arrive.noinc: the arrival fires when they have landed), thread 0 adds one ordinary
    // (releasing) arrival after writing the unit's descriptor word.  Copies of up to S units are in flight per thread.
    auto publish_unit = [&](uint32_t meta_word) {
      cp_async_mbar_arrive_noinc(&ctl->full[st_i]);
      if (ptid == 0) {
        ctl->meta[st_i] = meta_word;
        mbar_arrive(&ctl->full[st_i]);
      }
      ++gi;
      if (++st_i == S) { st_i = 0; st_ph ^= 1; }
    };
    auto acquire_stage = [&]() {        // the MMAs that read this stage one ring turn ago have completed
      if (gi >= (uint32_t)S) mbar_wait(&ctl->empty[st_i], st_ph ^ 1);
    };
    uint32_t tcount = 0;
    long long item = blockIdx.x;
    if (item < n_items) {
      fetch_idx(item, 0);
      cp_async_commit();
    }
    for (; item < n_items; item += gridDim.x, ++tcount) {
      const int buf = tcount & 1, slot = tcount % 3;
      const int32_t* my_idx = idx_s + (size_t)buf * kv * kWsRows + ptid;
      const bool row_ok = (item / n_ntiles) * kWsRows + ptid < n_out;
      // this tile's rulebook slice was committed (as a copy group of its own) one whole tile ago
      cp_async_wait<0>();
      // ---- which offsets have a partner anywhere in this tile ----
      for (int k = 0; k < kv; ++k) {
        const bool v = row_ok && my_idx[k * kWsRows] >= 0;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, v);
        if (lane == 0 && bal) atomicOr(&ctl->mask_x[slot][k >> 5], 1u << (k & 31));
      }
      if (ptid < 4) ctl->mask_x[(tcount + 1) % 3][ptid] = 0;
      named_bar_sync(1, 128);
      uint32_t mw[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) mw[w] = ld_shared_volatile_u32(&ctl->mask_x[slot][w]);
      int n_act = __popc(mw[0]) + __popc(mw[1]) + __popc(mw[2]) + __popc(mw[3]);
      uint8_t* act = act_s + buf * kWsMaxKV;
      {
        const int wq = ptid >> 5;      // ptid < 128: word of this thread's offset
        const uint32_t mine = wq == 0 ? mw[0] : (wq == 1 ? mw[1] : (wq == 2 ? mw[2] : mw[3]));
        if (ptid < kv && ((mine >> (ptid & 31)) & 1u)) {
          int pos = __popc(mine & ((1u << (ptid & 31)) - 1u));
          if (wq > 0) pos += __popc(mw[0]);
          if (wq > 1) pos += __popc(mw[1]);
          if (wq > 2) pos += __popc(mw[2]);
          act[pos] = (uint8_t)ptid;
        }
      }
      if (n_act == 0) { if (ptid == 0) act[0] = 0; n_act = 1; }   // a tile always carries at least one (all-zero) unit
      named_bar_sync(1, 128);
      // ---- next tile's rulebook slice: its own copy group (together with the previous tile's, long finished, row copies) ----
      if (item + gridDim.x < n_items) fetch_idx(item + gridDim.x, buf ^ 1);
      cp_async_commit();
      const int n0 = (int)(item % n_ntiles) * n_tile;
      for (int a = 0; a < n_act; ++a) {
        const int kidx = act[a];
        const int32_t src = row_ok ? my_idx[kidx * kWsRows] : -1;
        const T* g = feat + (int64_t)(src >= 0 ? src : 0) * c_in;
        for (int cc = 0; cc < n_cc; ++cc) {
          acquire_stage();
          uint8_t* a_s = ring + (size_t)st_i * stage_bytes;
          const uint32_t a_dst = smem_u32(a_s) + ptid * 16;
#pragma unroll
          for (int p = 0; p < LPR; ++p) cp_async16(a_dst + p * (kWsRows * 16), g + cc * KC + p * 8, src >= 0);
          if (!resident) load_w_tile(a_s + a_bytes, kidx, cc, n0, ptid, 128);
          publish_unit((uint32_t)kidx | ((uint32_t)cc << 10));
        }
      }
      // END marker: the MMA thread hands the accumulator to the epilogue warps
      acquire_stage();
      publish_unit(kMetaEnd);
    }
  } else if (warp == 8) {
    // ===================================== MMA issuer (whole warp waits, lane 0 issues) =====================================
    const uint32_t idesc = make_idesc(128, n_tile, UmmaFmt<T>::v, UmmaFmt<T>::v, 0, transpose_w ? 1 : 0);
    int st_i = 0, st_ph = 0;
    uint32_t tcount = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1, ause = tcount >> 1;
      const uint32_t d_tmem = tmem_base + buf * (cfg.tmem_cols >> 1);
      bool first = true;
      while (true) {
        const int s = st_i;
        mbar_wait(&ctl->full[s], st_ph);
        if (++st_i == S) { st_i = 0; st_ph ^= 1; }
        const uint32_t m = ld_shared_volatile_u32(&ctl->meta[s]);
        if (m & kMetaEnd) {
          if (lane == 0) {
            mma_commit(&ctl->acc_full[buf]);
            mbar_arrive(&ctl->empty[s]);
          }
          __syncwarp();
          break;
        }
        if (first && ause > 0) mbar_wait(&ctl->acc_empty[buf], (ause - 1) & 1);   // epilogue drained this buffer
        if (lane == 0) {
          fence_proxy_async();     // the stage was filled by cp.async (generic proxy); tcgen05.mma reads it through the async proxy
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)s * stage_bytes);
          const int kidx = m & 0x3FF, cc = (m >> 10) & 0x3FF;
          const uint32_t b_addr = resident ? smem_u32(w_res + (size_t)(kidx * n_cc + cc) * b_bytes) : a_addr + a_bytes;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t da = make_smem_desc(a_addr + 2 * ks * (kWsRows * 16), kWsRows * 16, 128);
            const uint64_t db = transpose_w ? make_smem_desc(b_addr + ks * 256, 128, KC * 16)
                                            : make_smem_desc(b_addr + 2 * ks * (n_tile * 16), n_tile * 16, 128);
            mma_ss(d_tmem, da, db, idesc, (first && ks == 0) ? 0u : 1u);
          }
          mma_commit(&ctl->empty[s]);
        }
        __syncwarp();
        first = false;
      }
    }
  } else {
    // ===================================== epilogue warps (TMEM lane quarter = warp) =====================================
    uint32_t tcount = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1, ause = tcount >> 1;
      mbar_wait(&ctl->acc_full[buf], ause & 1);
      tc_fence_after();
      const int64_t j = (item / n_ntiles) * kWsRows + warp * 32 + lane;
      const int n0 = (int)(item % n_ntiles) * n_tile;
      const uint32_t t_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (cfg.tmem_cols >> 1);
      for (int cb = 0; cb < n_tile; cb += 16) {
        uint32_t r[16];
        tmem_ld16(t_addr + cb, r);
        tmem_ld_wait();
        if (j < n_out) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
            if (bias) { v0 += to_f32(bias[n0 + cb + 2 * i]); v1 += to_f32(bias[n0 + cb + 2 * i + 1]); }
            w[i] = pack2<T>(v0, v1);
          }
          uint4* dst = reinterpret_cast<uint4*>(out + j * c_out + n0 + cb);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, cfg.tmem_cols);
}

template <typename T, int KC, int LAG>
inline void launch_conv_ws_kc(const ConvWsCfg& c, const void* feat, const void* weight, const void* bias, const int32_t* pair,
                              int64_t pair_stride, int64_t n_out, int c_in, int c_out, int kv, int transpose_w, int flip, void* out,
                              cudaStream_t stream) {
  cudaFuncSetAttribute(conv_ws_kernel<T, KC, LAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes);
  conv_ws_kernel<T, KC, LAG><<<c.grid, kWsThreads, c.smem_bytes, stream>>>((const T*)feat, (const T*)weight, (const T*)bias, pair, pair_stride,
                                                                           n_out, c_in, c_out, kv, transpose_w, flip, (T*)out, c);
}

template <typename T>
inline int launch_conv_ws_t(const void* feat, const void* weight, const void* bias, const int32_t* pair, int64_t pair_stride,
                            int64_t n_out, int c_in, int c_out, int kv, int transpose_w, int flip, void* out, cudaStream_t stream) {
  const ConvWsCfg c = conv_ws_cfg(n_out, c_in, c_out, kv);
#define B2PC_WS_GO(KC_, LAG_) launch_conv_ws_kc<T, KC_, LAG_>(c, feat, weight, bias, pair, pair_stride, n_out, c_in, c_out, kv, transpose_w, flip, out, stream)
  if (c.kc == 64) B2PC_WS_GO(64, 0); else if (c.kc == 32) B2PC_WS_GO(32, 0); else B2PC_WS_GO(16, 0);
#undef B2PC_WS_GO
  count_launches(1);
  B2PC_CHECK_LAUNCH("spconv_gather_gemm(tcgen05, warp-specialised)");
  return B2PC_OK;
}

inline int launch_conv_ws(const void* feat, const void* weight, const void* bias, const int32_t* pair, int64_t pair_stride,
                          int64_t n_out, int c_in, int c_out, int kv, int transpose_w, int flip, int dtype, void* out,
                          cudaStream_t stream) {
  if (n_out == 0) return B2PC_OK;
  if (dtype == B2PC_BF16)
    return launch_conv_ws_t<__nv_bfloat16>(feat, weight, bias, pair, pair_stride, n_out, c_in, c_out, kv, transpose_w, flip, out, stream);
  return launch_conv_ws_t<__half>(feat, weight, bias, pair, pair_stride, n_out, c_in, c_out, kv, transpose_w, flip, out, stream);
}

// =====================================================================================================================
// Weight gradient.  GEMM view: dW^T[(k, ci), co] = sum_j G[j, (k, ci)] * dout[j, co] with G the im2col rows
// G[j, (k, ci)] = feat[pair[k, j], ci].  The (k, ci) axis is cut into "slots" of mc channels (mc = largest power of two
// <= 128 dividing c_in); 128 / mc slots form one M tile (= 128 TMEM lanes), co is the N axis, rulebook rows are the
// reduction axis, consumed in chunks of 64 rows.  A CTA owns a group of M tiles (as many [128 x n_tile] fp32 accumulators
// as fit in the 512 TMEM columns) and one N tile, and sweeps its share of the row chunks: per chunk the dout slice is
// staged once (operand B, MN-major) and every M tile of the group gets one gathered operand-A tile (MN-major) and
// 4 tcgen05.mma (K = 16 rows each).  Row-range splits are reduced afterwards in a fixed order.
constexpr int kWg2Rows = 64;          // rulebook rows per chunk
constexpr int kWg2Threads = 160;      // warps 0-3 producers (+ final epilogue), warp 4 MMA
constexpr int kWg2MaxStages = 12;
constexpr int kWg2NB = 3;             // dout tile buffers

struct WgradWsCfg {
  int mc, n_chunks_c, spm, m_tiles, n_tile, n_ntiles, ncol, tpg, n_mgroups, n_splits, stages, lag, tmem_cols, a_bytes, b_bytes, idx_bytes,
      kspan, smem_bytes;
  long long n_row_chunks;
};

inline bool wgrad_ws_supported(int dtype, int c_in, int c_out, int kv) {
  if (dtype != B2PC_F16 && dtype != B2PC_BF16) return false;
  if (c_in % 16 != 0 || c_out % 16 != 0) return false;
  if (kv > kWsMaxKV) return false;
  return true;
}

inline WgradWsCfg wgrad_ws_cfg(int64_t n_out, int c_in, int c_out, int kv) {
  WgradWsCfg c;
  c.mc = 16;
  for (int m = 128; m >= 16; m >>= 1)
    if (c_in % m == 0) { c.mc = m; break; }
  c.n_chunks_c = c_in / c.mc;
  c.spm = 128 / c.mc;
  const int slots = kv * c.n_chunks_c;
  c.m_tiles = (slots + c.spm - 1) / c.spm;
  c.n_tile = 16;
  for (int nt = c_out <= 256 ? c_out : 256; nt >= 16; nt -= 16)
    if (c_out % nt == 0) { c.n_tile = nt; break; }
  c.n_ntiles = c_out / c.n_tile;
  c.ncol = 32;                          // accumulator slices sit on power-of-two column strides
  while (c.ncol < c.n_tile) c.ncol <<= 1;
  c.tpg = 512 / c.ncol;
  if (c.tpg > c.m_tiles) c.tpg = c.m_tiles;
  c.n_mgroups = (c.m_tiles + c.tpg - 1) / c.tpg;
  c.tmem_cols = 32;
  while (c.tmem_cols < c.tpg * c.ncol) c.tmem_cols <<= 1;
  c.n_row_chunks = ceil_div(n_out > 0 ? n_out : 1, kWg2Rows);
  const int groups = c.n_mgroups * c.n_ntiles;
  // one persistent CTA per SM (the deep gather ring hides the latency): the row range is split so that groups x splits fills the
  // SMs in ONE wave -- rounding up (162 CTAs for 27 groups, 154 for 7, 216 for 108) left a second, nearly empty wave that doubled
  // the runtime of the deep levels
  long long sp = kNumSMs / groups;
  if (sp > c.n_row_chunks) sp = c.n_row_chunks;
  if (sp < 1) sp = 1;
  c.n_splits = (int)sp;
  c.a_bytes = 128 * kWg2Rows * 2;
  c.b_bytes = kWg2Rows * c.n_tile * 2;
  // kernel offsets one CTA touches: the slots of its M-tile group span at most this many offsets
  c.kspan = (c.tpg * c.spm + c.n_chunks_c - 1) / c.n_chunks_c + 1;
  if (c.kspan > kv) c.kspan = kv;
  c.idx_bytes = 2 * c.kspan * kWg2Rows * 4;
  int st = (kWsSmemBudget - kWsCtrlBytes - kWg2NB * c.b_bytes - c.idx_bytes) / c.a_bytes;
  if (st > kWg2MaxStages) st = kWg2MaxStages;
  if (st < 3) st = 3;
  c.stages = st;
  c.lag = st >= 8 ? 6 : 2;
  c.smem_bytes = kWsCtrlBytes + c.idx_bytes + kWg2NB * c.b_bytes + c.stages * c.a_bytes;
  return c;
}

inline size_t wgrad_ws_workspace_bytes(int64_t n_out, int c_in, int c_out, int kv) {
  if (c_in % 16 != 0 || c_out % 16 != 0) return 0;
  const WgradWsCfg c = wgrad_ws_cfg(n_out, c_in, c_out, kv);
  return c.n_splits > 1 ? (size_t)c.n_splits * c_out * kv * c_in * sizeof(float) + 256 : 256;
}

struct WgCtrl {
  uint64_t full[kWg2MaxStages];
  uint64_t empty[kWg2MaxStages];
  uint64_t bempty[kWg2NB];
  uint64_t done;
  uint32_t tmem_slot;
};
static_assert(sizeof(WgCtrl) <= kWsCtrlBytes, "control block too large");

// Producer mapping: thread (r = tid & 63, h = tid >> 6) owns row r of the 64-row chunk and, of every slot of a unit, the h-th half
// of its MC channels (MC/16 contiguous 16-byte pieces from one base pointer).  The rulebook slice of the NEXT chunk (only the
// kernel offsets this CTA's M-tile group touches) is fetched by 4-byte cp.async into a second buffer while the current chunk streams.
template <typename T, int MC, int LAG>
__global__ void __launch_bounds__(kWg2Threads)
wgrad_ws_kernel(const T* __restrict__ feat, const T* __restrict__ dout, const int32_t* __restrict__ pair, int64_t pair_stride,
                int64_t n_out, int c_in, int c_out, int kv, float* __restrict__ dst_base, WgradWsCfg cfg) {
  using namespace umma;
  extern __shared__ __align__(128) uint8_t smem[];
  WgCtrl* ctl = reinterpret_cast<WgCtrl*>(smem);
  int32_t* idx_s = reinterpret_cast<int32_t*>(smem + kWsCtrlBytes);       // [2][kspan][64]
  uint8_t* b_ring = smem + kWsCtrlBytes + cfg.idx_bytes;
  uint8_t* a_ring = b_ring + kWg2NB * cfg.b_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int SPM = 128 / MC;     // slots per M tile
  constexpr int LPR = MC / 8;       // 16-byte pieces per gathered row of one slot
  constexpr int PPT = LPR / 2;      // pieces per producer thread per slot (two threads share a row)
  const int S = cfg.stages, n_tile = cfg.n_tile;
  const int split = blockIdx.x, n_splits = gridDim.x;
  const int mgroup = blockIdx.y % cfg.n_mgroups, nt_i = blockIdx.y / cfg.n_mgroups;
  const int mt0 = mgroup * cfg.tpg;
  const int n_mt = min(cfg.tpg, cfg.m_tiles - mt0);      // M tiles of this CTA
  const int n0 = nt_i * n_tile;
  const int n_cc = cfg.n_chunks_c;
  const int slots_total = kv * n_cc;
  const long long n_rc = cfg.n_row_chunks;
  const int a_bytes = cfg.a_bytes, b_bytes = cfg.b_bytes, kspan = cfg.kspan;
  const int k_lo = (mt0 * SPM) / n_cc;                   // first kernel offset this CTA touches
  const int k_cnt = min(kspan, kv - k_lo);

  if (warp == 0) { tmem_alloc(&ctl->tmem_slot, cfg.tmem_cols); tmem_relinquish(); }
  if (tid == 32) {
    for (int s = 0; s < S; ++s) { mbar_init(&ctl->full[s], 128); mbar_init(&ctl->empty[s], 1); }
    for (int b = 0; b < kWg2NB; ++b) mbar_init(&ctl->bempty[b], 1);
    mbar_init(&ctl->done, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_volatile_u32(&ctl->tmem_slot);

  if (warp < 4) {
    // ===================================== producers =====================================
    const int r = tid & 63, h = tid >> 6;
    auto fetch_idx = [&](long long rc, int buf) {   // thread (r, h) fetches row r of the offsets k_lo + h, k_lo + h + 2, ...
      int64_t j = rc * kWg2Rows + r;
      if (j >= n_out) j = n_out - 1;
      const uint32_t dst = smem_u32(idx_s + (size_t)buf * kspan * kWg2Rows + r);
      for (int kk = h; kk < k_cnt; kk += 2) cp_async4(dst + kk * (kWg2Rows * 4), pair + (int64_t)(k_lo + kk) * pair_stride + j, true);
    };
    uint32_t gi = 0;
    int st_i = 0, st_ph = 0;
    uint32_t chunk_local = 0;
    long long rc = split;
    if (rc < n_rc) {
      fetch_idx(rc, 0);
      cp_async_commit();
    }
    for (; rc < n_rc; rc += n_splits, ++chunk_local) {
      const int buf = chunk_local & 1;
      cp_async_wait<0>();                          // this chunk's rulebook slice (own copy group, committed one chunk ago) has landed
      named_bar_sync(1, 128);                      // ... and so has the half fetched by the other thread of this row
      if (rc + n_splits < n_rc) fetch_idx(rc + n_splits, buf ^ 1);
      cp_async_commit();
      const bool row_ok = rc * kWg2Rows + r < n_out;
      const int32_t* my_idx = idx_s + (size_t)buf * kspan * kWg2Rows + r;
      for (int mt = 0; mt < n_mt; ++mt) {
        if (gi >= (uint32_t)S) mbar_wait(&ctl->empty[st_i], st_ph ^ 1);
        if (mt == 0) {   // dout tile of this chunk (operand B, MN-major planes): tracked by the first unit's barrier
          const int b = chunk_local % kWg2NB;
          if (chunk_local >= (uint32_t)kWg2NB) mbar_wait(&ctl->bempty[b], ((chunk_local / kWg2NB) - 1) & 1);
          const int ppr = n_tile / 8;
          const uint32_t b_dst = smem_u32(b_ring + (size_t)b * b_bytes);
          for (int q = tid; q < kWg2Rows * ppr; q += 128) {
            const int rr = q / ppr, p = q % ppr;
            const int64_t j = rc * kWg2Rows + rr;
            cp_async16(b_dst + p * (kWg2Rows * 16) + rr * 16, dout + (j < n_out ? j : 0) * c_out + n0 + p * 8, j < n_out);
          }
        }
        const uint32_t a_dst = smem_u32(a_ring + (size_t)st_i * a_bytes) + r * 16;
        int slot = (mt0 + mt) * SPM;
        int k = slot / n_cc, cchunk = slot % n_cc;
#pragma unroll
        for (int sl = 0; sl < SPM; ++sl) {
          const int32_t src = (slot < slots_total && row_ok) ? my_idx[(k - k_lo) * kWg2Rows] : -1;
          const T* g = feat + (int64_t)(src >= 0 ? src : 0) * c_in + cchunk * MC + h * (PPT * 8);
#pragma unroll
          for (int q = 0; q < PPT; ++q) cp_async16(a_dst + (sl * LPR + h * PPT + q) * (kWg2Rows * 16), g + q * 8, src >= 0);
          ++slot;
          if (++cchunk == n_cc) { cchunk = 0; ++k; }
        }
        // hand-over without a wait: the stage's barrier tracks this thread's outstanding copies (arrival fires when they landed)
        cp_async_mbar_arrive_noinc(&ctl->full[st_i]);
        ++gi;
        if (++st_i == S) { st_i = 0; st_ph ^= 1; }
      }
    }
    // ===================================== epilogue (same warps; TMEM lane quarter = warp) =====================================
    mbar_wait(&ctl->done, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;            // accumulator row = (slot in tile, channel in slot)
    const int sl = m / MC, ch = m % MC;
    const int64_t kvc = (int64_t)kv * c_in;
    float* dst_split = dst_base + (int64_t)split * c_out * kvc;
    for (int t = 0; t < n_mt; ++t) {
      const int slot = (mt0 + t) * SPM + sl;
      const bool ok = slot < slots_total;
      const int k = slot / n_cc, ci = (slot % n_cc) * MC + ch;
      const uint32_t t_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + t * cfg.ncol;
      for (int cb = 0; cb < n_tile; cb += 16) {
        uint32_t rg[16];
        tmem_ld16(t_addr + cb, rg);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int e = 0; e < 16; ++e) dst_split[(int64_t)(n0 + cb + e) * kvc + (int64_t)k * c_in + ci] = __uint_as_float(rg[e]);
        }
      }
    }
  } else {
    // ===================================== MMA issuer (whole warp waits, lane 0 issues) =====================================
    const uint32_t idesc = make_idesc(128, n_tile, UmmaFmt<T>::v, UmmaFmt<T>::v, 1, 1);
    int st_i = 0, st_ph = 0;
    uint32_t chunk_local = 0;
    for (long long rc = split; rc < n_rc; rc += n_splits, ++chunk_local) {
      const int b = chunk_local % kWg2NB;
      const uint32_t b_addr = smem_u32(b_ring + (size_t)b * b_bytes);
      for (int t = 0; t < n_mt; ++t) {
        const int s = st_i;
        mbar_wait(&ctl->full[s], st_ph);
        if (++st_i == S) { st_i = 0; st_ph ^= 1; }
        if (lane == 0) {
          fence_proxy_async();     // operands were written by cp.async (generic proxy), tcgen05.mma reads through the async proxy
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_ring + (size_t)s * a_bytes);
#pragma unroll
          for (int ks = 0; ks < kWg2Rows / 16; ++ks)
            mma_ss(tmem_base + t * cfg.ncol, make_smem_desc(a_addr + ks * 256, 128, kWg2Rows * 16),
                   make_smem_desc(b_addr + ks * 256, 128, kWg2Rows * 16), idesc, (chunk_local > 0 || ks > 0) ? 1u : 0u);
          mma_commit(&ctl->empty[s]);
          if (t == n_mt - 1) mma_commit(&ctl->bempty[b]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) mma_commit(&ctl->done);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, cfg.tmem_cols);
}

template <typename T, int MC, int LAG>
inline void launch_wgrad_ws_mc(const WgradWsCfg& c, const void* feat, const void* dout, const int32_t* pair, int64_t pair_stride,
                               int64_t n_out, int c_in, int c_out, int kv, float* dst, cudaStream_t stream) {
  cudaFuncSetAttribute(wgrad_ws_kernel<T, MC, LAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes);
  dim3 grid(c.n_splits, c.n_mgroups * c.n_ntiles);
  wgrad_ws_kernel<T, MC, LAG><<<grid, kWg2Threads, c.smem_bytes, stream>>>((const T*)feat, (const T*)dout, pair, pair_stride, n_out, c_in,
                                                                          c_out, kv, dst, c);
}

template <typename T>
inline int launch_wgrad_ws_t(const void* feat, const void* dout, const int32_t* pair, int64_t pair_stride, int64_t n_out, int c_in,
                             int c_out, int kv, float* dweight, void* ws, cudaStream_t stream) {
  const WgradWsCfg c = wgrad_ws_cfg(n_out, c_in, c_out, kv);
  float* dst = c.n_splits > 1 ? (float*)ws : dweight;
#define B2PC_WG_GO(MC_, LAG_) launch_wgrad_ws_mc<T, MC_, LAG_>(c, feat, dout, pair, pair_stride, n_out, c_in, c_out, kv, dst, stream)
  switch (c.mc) { case 128: B2PC_WG_GO(128, 0); break; case 64: B2PC_WG_GO(64, 0); break; case 32: B2PC_WG_GO(32, 0); break; default: B2PC_WG_GO(16, 0); }
#undef B2PC_WG_GO
  count_launches(1);
  if (c.n_splits > 1) {
    const int64_t elems = (int64_t)c_out * kv * c_in;
    int rb = (int)ceil_div(elems, 256);
    if (rb > kNumSMs * 8) rb = kNumSMs * 8;
    reduce_splits_kernel<<<rb, 256, 0, stream>>>((const float*)ws, elems, c.n_splits, dweight);
    count_launches(1);
  }
  B2PC_CHECK_LAUNCH("spconv_bwd_weight(tcgen05, warp-specialised)");
  return B2PC_OK;
}

inline int launch_wgrad_ws(const void* feat, const void* dout, const int32_t* pair, int64_t pair_stride, int64_t n_out, int c_in,
                           int c_out, int kv, int dtype, float* dweight, void* ws, cudaStream_t stream) {
  if (dtype == B2PC_BF16)
    return launch_wgrad_ws_t<__nv_bfloat16>(feat, dout, pair, pair_stride, n_out, c_in, c_out, kv, dweight, ws, stream);
  return launch_wgrad_ws_t<__half>(feat, dout, pair, pair_stride, n_out, c_in, c_out, kv, dweight, ws, stream);
}

}  // namespace b2pc
