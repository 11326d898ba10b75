// Sparse convolution as an output-stationary implicit GEMM on tcgen05 tensor cores.
//
//   out[j, :] = bias + sum_k  feat[pair[k', j], :] @ W_k            (k' = flip ? KV-1-k : k)
//
// One CTA owns 128 consecutive output rows x n_tile output channels; the fp32 accumulator [128 x n_tile] lives in
// TMEM for the whole sweep over kernel offsets.  Per (active offset k, channel chunk of KC): thread t gathers row
// pair[k', row0+t] (KC*2 contiguous bytes, zero-filled when the pair is absent) into a K-major operand tile with
// cp.async, the matching W_k slice is staged next to it, and one thread issues KC/16 tcgen05.mma (M=128, N=n_tile,
// K=16).  Offsets with no partner in the whole tile are skipped.  A ring of stages keeps several gathers in flight
// while the tensor core drains earlier ones; stage reuse is gated by tcgen05.commit -> mbarrier.
// Every output element is written once.  When the launch is under-filled (deep levels) the offsets are split over several CTAs per
// tile; each split stores its fp32 partial tile to its own slice of a scratch buffer and a small kernel sums the slices in a fixed
// order (deterministic; the round-1 red.global.add version was order-dependent and bound by L2 atomic throughput).
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "attn_umma.cuh"  // UmmaFmt, pack2
#include <cstring>
#include <type_traits>

namespace b2pc {

constexpr int kCuM = 128;       // output rows per CTA (= threads)
constexpr int kCuMaxKV = 32;    // kernel volume limit of this path (27 for 3^3, 8 for 2^3)
constexpr int kCuMaxStages = 4;   // ring depth (prefetch distance 2)

struct ConvUmmaCfg { int kc, n_tile, tmem_cols, smem_bytes, idx_rows, ksplit; };

inline ConvUmmaCfg conv_umma_cfg(int64_t n_out, int c_in, int c_out, int kv) {
  ConvUmmaCfg c;
  c.kc = c_in % 64 == 0 ? 64 : (c_in % 32 == 0 ? 32 : 16);
  // widest legal N tile (A rows are gathered once per N tile); when that leaves the GPU under-filled (deep, narrow-N
  // levels: a few dozen row tiles, 27 offsets x C/64 chunks of strictly sequential work each) the kernel offsets are split
  // over `ksplit` CTAs per tile, which store their partial sums to per-split slices of an fp32 scratch (summed afterwards)
  const int64_t m_tiles = ceil_div(n_out > 0 ? n_out : 1, kCuM);
  // A/B knobs (tools/probe_conv.py sweeps them inside one process when B2PC_CONV_TUNE=1 was set at start-up; otherwise read once)
  static const bool tune = [] { const char* e = getenv("B2PC_CONV_TUNE"); return e && atoi(e) != 0; }();
  auto knob = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
  static const int nt0 = knob("B2PC_CONV_NT"), ks0 = knob("B2PC_CONV_KSPLIT"), ksx0 = knob("B2PC_CONV_KS");
  const int env_nt = tune ? knob("B2PC_CONV_NT") : nt0;        // cap of the N tile (under-filled levels only)
  const int env_ks = tune ? knob("B2PC_CONV_KSPLIT") : ks0;    // cap of the offset split
  const int env_ksx = tune ? knob("B2PC_CONV_KS") : ksx0;      // exact offset split
  int nt_cap = c_out <= 256 ? c_out : 256;
  if (env_nt > 0 && m_tiles < 100 && nt_cap > env_nt) nt_cap = env_nt;   // only the under-filled (deep) levels are affected
  c.n_tile = 16;
  for (int nt = nt_cap; nt >= 16; nt -= 16)
    if (c_out % nt == 0) { c.n_tile = nt; break; }
  c.tmem_cols = 32;
  while (c.tmem_cols < c.n_tile) c.tmem_cols <<= 1;
  const int unit_bytes = kCuM * c.kc * 2 + c.n_tile * c.kc * 2;   // one ring stage: A tile + B tile of one (offset, chunk) unit
  c.idx_rows = kv < kCuMaxKV ? kv : kCuMaxKV;
  c.smem_bytes = c.idx_rows * kCuM * 4 + 256 + kCuMaxStages * unit_bytes;
  // offset split of under-filled launches: CTAs per SM from shared memory / TMEM, then the split that minimises
  //   waves(ctas * ks) * (offsets per CTA + fixed prologue/epilogue) + cost of summing ks partial tiles
  // -- in particular never a split that spills a few CTAs into a second wave (30 tiles x 5 = 150 CTAs on 148 SMs)
  const int64_t ctas = m_tiles * (c_out / c.n_tile);
  int occ = (227 * 1024) / (c.smem_bytes + 1024);
  if (occ > 512 / c.tmem_cols) occ = 512 / c.tmem_cols;
  if (occ < 1) occ = 1;
  const int64_t slots = (int64_t)kNumSMs * occ;
  c.ksplit = 1;
  if (ctas < slots && kv > 1) {
    const int cap = env_ks > 0 ? env_ks : 16;
    double best = 1e30;
    for (int ks = 1; ks <= cap && ks <= kv; ++ks) {
      const double waves = (double)ceil_div(ctas * ks, slots);
      const double cost = waves * ((double)ceil_div(kv, ks) + 3.0) + 0.3 * ks;
      if (cost < best - 1e-9) { best = cost; c.ksplit = ks; }
    }
    if (env_ksx > 0) c.ksplit = env_ksx < kv ? env_ksx : kv;
  }
  return c;
}

inline bool spconv_umma_supported(int dtype, int c_in, int c_out) {
  if (dtype != B2PC_F16 && dtype != B2PC_BF16) return false;
  if (c_in % 16 != 0 || c_out % 16 != 0) return false;
  if (c_out > 256 && c_out % 64 != 0) return false;
  return true;
}

// tile::gather4 maps of the feature matrix, one per channel chunk: map i views columns [i*kc, (i+1)*kc) (base pointer advanced, row
// stride unchanged) so that the gather always reads from column 0.  (A single map with a non-zero column coordinate returned wrong
// data on B200 for every chunk but the first -- measured, tools/conv_ta_check.py; the box of a gather4 map is {kc, 1}: four rows
// in the box is an illegal instruction.)
constexpr int kCuMaxChunks = 8;
struct ConvGatherMaps { CUtensorMap m[kCuMaxChunks]; };

// KC_T / NT_T: compile-time channel chunk and N tile for the common layer widths (0 = use the runtime arguments); NCC_T: compile-time
// number of channel chunks per offset (0 = runtime).  The specialisations fold the address arithmetic of the staging loops; the
// ring depth is a compile-time constant and the (offset, chunk) cursor advances incrementally, so the per-iteration instruction
// stream holds no integer division (ncu of the round-1 kernel: ~250 warp instructions per iteration, most of them index math).
// MB: hand the staged tiles to the MMA-issuing thread through a per-stage `full` mbarrier (every thread's cp.async.mbarrier.arrive.noinc
// fires when its copies of that stage have landed) instead of cp.async.wait_group + a block-wide barrier per iteration; only thread 0
// waits, the other warps run ahead by up to the ring depth.
// TB: the weight tile of a unit is ONE swizzled TMA box (rows of kc channels = the swizzle span) issued by one thread and tracked by a
// per-stage byte-counting mbarrier, instead of 1-16 cp.async per thread: cp.async issue is the limiter of this kernel (each 16-byte
// warp instruction costs ~30-50 cycles of LSU time), and at C >= 128 the weights are two thirds of those instructions.  K-major only:
// the backward-data pass runs on a transposed copy of the weights (conv_transpose_w_kernel) so that both passes take this path.
// TA (with TB): the gathered rows come by TMA too -- tile::gather4, four rulebook rows per instruction, one instruction per lane of a
// producer warp per unit, absent pairs (-1) as out-of-range rows that the TMA zero-fills; the tile is then a swizzled K-major operand
// like the weights.  No thread issues cp.async any more: warp 1 produces, thread 0 issues the MMAs, both synchronise through the
// per-stage mbarriers only (no block barrier, no wait_group in the loop).  Measured on B200 (profiles/README.md): bit-correct, but the
// TMA unit retires only about one gather4 per 44 cycles per SM (C = 32: 0.103 -> 0.247 ms, C = 64: 0.198 -> 0.292 ms), so cp.async with
// thread = row stays the default for the gathered operand and B2PC_CONV_TMAA=1 selects this path.
template <typename T, int KC_T, int NT_T, int NCC_T, bool MB, bool TB, bool TA>
__global__ void __launch_bounds__(kCuM)
gather_gemm_umma_kernel(const T* __restrict__ feat, const T* __restrict__ weight, const T* __restrict__ bias,
                        const int32_t* __restrict__ pair, int64_t pair_stride, int64_t n_out, int c_in, int c_out, int kv,
                        int transpose_w, int flip, T* __restrict__ out, int kc_arg, int n_tile_arg, int tmem_cols, int idx_rows,
                        int ksplit, float* __restrict__ acc, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ ConvGatherMaps tmap_f,
                        int n_in) {
  using namespace umma;
  constexpr int S = kCuMaxStages;   // ring depth
  constexpr int PD = S - 2;         // prefetch distance: a refilled stage was consumed two iterations ago, so its MMAs are (almost
                                    // always) complete already and the stage-free wait does not serialise on the tensor pipe
  const int kc = KC_T ? KC_T : kc_arg;
  const int n_tile = NT_T ? NT_T : n_tile_arg;
  const int n_cc = NCC_T ? NCC_T : c_in / kc;
  extern __shared__ __align__(128) uint8_t smem[];
  int32_t* idx_s = reinterpret_cast<int32_t*>(smem);                       // [idx_rows][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + idx_rows * kCuM * 4);  // [S]
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(bars + S);
  uint32_t* tmem_slot = mask_s + 1;
  uint8_t* act_s = reinterpret_cast<uint8_t*>(tmem_slot + 1);   // [kCuMaxKV]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + idx_rows * kCuM * 4 + 128);   // [S] (MB only)
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + idx_rows * kCuM * 4 + 160);  // [S] (TB only): weight tile landed
  uint8_t* stage0 = smem + idx_rows * kCuM * 4 + 256;
  if (TB) stage0 += (1024u - (smem_u32(stage0) & 1023u)) & 1023u;   // swizzled tiles: 1024-byte aligned atoms
  const int a_bytes = kCuM * kc * 2, b_bytes = n_tile * kc * 2, stage_bytes = a_bytes + b_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kCuM;
  const int n0 = blockIdx.y * n_tile;

  if (warp == 0) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bars[s], 1); if (MB) mbar_init(&full[s], kCuM); if (TB) mbar_init(&wfull[s], 1); }
    if (TA) mbar_init(&full[0], 1);   // TA: `all MMAs done` (the idle warps cannot use the ring barriers: see below)
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc(128, n_tile, UmmaFmt<T>::v, UmmaFmt<T>::v, 0, transpose_w ? 1 : 0);
  const uint32_t stage0_u32 = smem_u32(stage0);

  // ---- loop-invariant staging addresses -------------------------------------------------------------------------------------
  // A: thread = row; kc/8 pieces of 16 B -> plane p at p*2048 + row*16.  (Letting the lanes of a warp sweep the pieces of a row first
  // -- whole row segments per instruction, 4-8 x fewer L1 tag wavefronts -- was measured SLOWER on B200: 0.110 -> 0.119 ms at C = 32,
  // 0.268 -> 0.336 ms at C = 64; the lanes then collide on the shared-memory banks of the plane layout.)
  const uint32_t a_dst0 = stage0_u32 + tid * 16;
  // B, K-major (forward): rows n (c_out side), kc contiguous channels; piece (n, p) -> p*(n_tile*16) + n*16; thread handles pieces
  // q = tid + i*128, i.e. p = tid % ppr (fixed) and n = tid / ppr + i * (128 / ppr)
  // B, MN-major (backward-data): rows kk (weight's c_out axis = reduction side), n_tile contiguous; piece (kk, p) -> p*(kc*16) + kk*16
  const int ppr = transpose_w ? n_tile / 8 : kc / 8;                 // pieces per weight row
  const int b_total = transpose_w ? kc * ppr : n_tile * ppr;         // pieces per B tile
  const int b_r = tid / ppr, b_p = tid - b_r * ppr;
  const int rows_per_pass = kCuM / ppr;                              // ppr <= 32 divides 128
  const T* b_src0;
  int64_t b_src_step;
  uint32_t b_dst0, b_dst_step;
  if (!transpose_w) {
    b_src0 = weight + (int64_t)(n0 + b_r) * kv * c_in + b_p * 8;
    b_src_step = (int64_t)rows_per_pass * kv * c_in;
    b_dst0 = stage0_u32 + a_bytes + b_p * (n_tile * 16) + b_r * 16;
  } else {
    b_src0 = weight + (int64_t)b_r * kv * c_out + n0 + b_p * 8;
    b_src_step = (int64_t)rows_per_pass * kv * c_out;
    b_dst0 = stage0_u32 + a_bytes + b_p * (kc * 16) + b_r * 16;
  }
  b_dst_step = rows_per_pass * 16;
  // UMMA descriptors of stage 0; other stages / K steps add a 16-byte-unit offset to the address field
  const uint64_t da0 = TA ? make_smem_desc_swz(stage0_u32, kc * 2) : make_smem_desc(stage0_u32, kCuM * 16, 128);
  const uint64_t db0 = TB ? make_smem_desc_swz(stage0_u32 + a_bytes, kc * 2)
                          : (transpose_w ? make_smem_desc(stage0_u32 + a_bytes, 128, kc * 16) : make_smem_desc(stage0_u32 + a_bytes, n_tile * 16, 128));
  const uint32_t da_ks = TA ? 2 : ((2 * kCuM * 16) >> 4), db_ks = TB ? 2 : ((transpose_w ? 256 : 2 * n_tile * 16) >> 4), d_stage = stage_bytes >> 4;

  int gi = 0;                 // iterations issued so far (uniform); stage = gi % S, use count = gi / S

  // kernel offsets are processed in chunks of kCuMaxKV (the rulebook slice of a chunk lives in shared memory)
  for (int kb = 0; kb < kv; kb += kCuMaxKV) {
    const int kcnt = min(kCuMaxKV, kv - kb);
    __syncthreads();
    if (tid == 0) *mask_s = 0;
    __syncthreads();
    {
      const int64_t j = row0 + tid;
      uint32_t wmask = 0;
      for (int k8 = 0; k8 < kcnt; k8 += 8) {
        int32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // 8 independent loads in flight before the first use
          const int k = k8 + u;
          const int kp = flip ? kv - 1 - (kb + k) : kb + k;
          v[u] = (k < kcnt && j < n_out) ? __ldg(pair + (int64_t)kp * pair_stride + j) : -1;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = k8 + u;
          if (k < kcnt) {
            idx_s[k * kCuM + tid] = v[u];
            if (__ballot_sync(0xFFFFFFFFu, v[u] >= 0)) wmask |= 1u << k;
          }
        }
      }
      if (lane == 0 && wmask) atomicOr(mask_s, wmask);
    }
    __syncthreads();
    uint32_t mask = *mask_s;
    if (ksplit > 1) {   // this CTA's share of the offsets: k % ksplit == blockIdx.z
      uint32_t mine = 0;
      for (int k = 0; k < kcnt; ++k)
        if ((kb + k) % ksplit == (int)blockIdx.z) mine |= 1u << k;
      mask &= mine;
    }
    const int n_act = __popc(mask);
    const int n_it = n_act * n_cc;                // iteration = unit = (active offset, channel chunk)
    if (tid < n_act) act_s[tid] = (uint8_t)__fns(mask, 0, tid + 1);   // active offsets of this chunk, in order
    __syncthreads();

    if (TA) {
      if (warp == 1) {          // ---- producer warp: one gather4 per lane (rows 4*lane .. 4*lane+3) + one weight box per unit
        int pa = 0, pc = 0;
        for (int it = 0; it < n_it; ++it) {
          const int g = gi + it, s = g & (S - 1);
          if (g >= S) mbar_wait(&bars[s], ((g / S) - 1) & 1);          // the MMAs that read this stage S units ago have completed
          const int kl = act_s[pa];
          const int c0 = pc * kc;   // weights: column of the [c_out, kv*c_in] matrix; features: one map per channel chunk (see ConvGatherMaps)
          const uint32_t st = stage0_u32 + s * stage_bytes;
          if (lane == 0) {
            mbar_expect_tx(&wfull[s], (uint32_t)(a_bytes + b_bytes));
            tma_load_2d(st + a_bytes, &tmap_w, (kb + kl) * c_in + c0, n0, &wfull[s]);
          }
          const int4 r4 = *reinterpret_cast<const int4*>(idx_s + kl * kCuM + 4 * lane);
          tma_gather4(st + lane * (4 * kc * 2), &tmap_f.m[pc], 0, r4.x >= 0 ? r4.x : n_in, r4.y >= 0 ? r4.y : n_in, r4.z >= 0 ? r4.z : n_in,
                      r4.w >= 0 ? r4.w : n_in, &wfull[s]);
          if (++pc == n_cc) { pc = 0; ++pa; }
        }
      } else if (tid == 0) {    // ---- MMA issuer
        for (int it = 0; it < n_it; ++it) {
          const int g = gi + it, s = g & (S - 1);
          mbar_wait(&wfull[s], (g / S) & 1);
          tc_fence_after();
          const uint64_t da = da0 + (uint64_t)(s * d_stage), db = db0 + (uint64_t)(s * d_stage);
#pragma unroll
          for (int ks = 0; ks < (KC_T ? KC_T / 16 : 4); ++ks)
            if (KC_T || ks < kc / 16) mma_ss(tmem_base, da + ks * da_ks, db + ks * db_ks, idesc, (g > 0 || ks > 0) ? 1u : 0u);
          mma_commit(&bars[s]);
        }
      }
      gi += n_it;
      continue;
    }
    int ld_a = 0, ld_cc = 0;   // cursor of the next unit to load: index into act_s, channel chunk
    auto issue_loads = [&](int s) {
      const int kl = act_s[ld_a];                     // offset inside the chunk
      const int k = kb + kl;                          // weight slice
      const int c0 = ld_cc * kc;
      const uint32_t s_off = s * stage_bytes;
      const int32_t src = idx_s[kl * kCuM + tid];
      const T* g = feat + (int64_t)(src >= 0 ? src : 0) * c_in + c0;
      const uint32_t a_dst = a_dst0 + s_off;
#pragma unroll
      for (int p = 0; p < (KC_T ? KC_T / 8 : 8); ++p)
        if (KC_T || p < kc / 8) cp_async16(a_dst + p * (kCuM * 16), g + p * 8, src >= 0);
      if (TB) {
        if (tid == 32) {   // one thread, one box: rows n0 .. n0+n_tile of the [c_out, kv*c_in] weight matrix, columns k*c_in + c0 .. +kc
          mbar_expect_tx(&wfull[s], (uint32_t)b_bytes);
          tma_load_2d(stage0_u32 + a_bytes + s_off, &tmap_w, k * c_in + c0, n0, &wfull[s]);
        }
        if (++ld_cc == n_cc) { ld_cc = 0; ++ld_a; }
        return;
      }
      const T* bs = b_src0 + (transpose_w ? ((int64_t)c0 * kv + k) * c_out : (int64_t)k * c_in + c0);
      uint32_t bd = b_dst0 + s_off;
      if (KC_T && NT_T) {
        constexpr int kPieces = (KC_T ? KC_T : 8) * (NT_T ? NT_T : 8) / 8;     // same count for both operand majors
        if (kPieces >= kCuM) {
#pragma unroll
          for (int i = 0; i < kPieces / kCuM; ++i) { cp_async16(bd, bs, true); bs += b_src_step; bd += b_dst_step; }
        } else if (tid < kPieces) {
          cp_async16(bd, bs, true);
        }
      } else {   // runtime tile shape (ppr need not divide 128): per-piece index arithmetic
        const T* wk = weight + (transpose_w ? ((int64_t)c0 * kv + k) * c_out + n0 : ((int64_t)n0 * kv + k) * c_in + c0);
        const int64_t row_stride = (int64_t)kv * (transpose_w ? c_out : c_in);
        const uint32_t plane = (transpose_w ? kc : n_tile) * 16;
        for (int q = tid; q < b_total; q += kCuM) {
          const int r = q / ppr, pp = q - r * ppr;
          cp_async16(stage0_u32 + a_bytes + s_off + pp * plane + r * 16, wk + r * row_stride + pp * 8, true);
        }
      }
      if (++ld_cc == n_cc) { ld_cc = 0; ++ld_a; }
    };
    auto stage_free = [&](int g) {   // stage g % S was last read by the MMAs of global iteration g - S
      if (g >= S) mbar_wait(&bars[g & (S - 1)], ((g / S) - 1) & 1);
    };

    auto issue_mma = [&](int g) {
      const int s = g & (S - 1);
      if (TB) mbar_wait(&wfull[s], (g / S) & 1);   // TMA writes through the async proxy: visible to the MMA once the bytes are counted
      tc_fence_after();
      const uint64_t da = da0 + (uint64_t)(s * d_stage), db = db0 + (uint64_t)(s * d_stage);
#pragma unroll
      for (int ks = 0; ks < (KC_T ? KC_T / 16 : 4); ++ks)
        if (KC_T || ks < kc / 16) mma_ss(tmem_base, da + ks * da_ks, db + ks * db_ks, idesc, (g > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&bars[s]);
    };
    if (MB) {
#pragma unroll
      for (int it = 0; it < PD; ++it)
        if (it < n_it) { stage_free(gi + it); issue_loads((gi + it) & (S - 1)); cp_async_mbar_arrive_noinc(&full[(gi + it) & (S - 1)]); }
      for (int it = 0; it < n_it; ++it) {
        if (it + PD < n_it) {
          const int g2 = gi + it + PD;
          stage_free(g2);
          issue_loads(g2 & (S - 1));
          cp_async_mbar_arrive_noinc(&full[g2 & (S - 1)]);
        }
        if (tid == 0) {
          const int g = gi + it;
          mbar_wait(&full[g & (S - 1)], (g / S) & 1);
          fence_proxy_async();
          issue_mma(g);
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < PD; ++it) {
        if (it < n_it) { stage_free(gi + it); issue_loads((gi + it) & (S - 1)); }
        cp_async_commit();
      }
      for (int it = 0; it < n_it; ++it) {
        if (it + PD < n_it) { stage_free(gi + it + PD); issue_loads((gi + it + PD) & (S - 1)); }
        cp_async_commit();
        cp_async_wait<PD>();   // groups committed: PD + it + 1; iteration `it` is group #it -> PD may stay pending
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) issue_mma(gi + it);
      }
    }
    gi += n_it;
    cp_async_wait<0>();
  }
  const int n_it = gi;
  if (TA) {
    // the warps that neither produce nor issue arrive here at once; a parity wait on a ring barrier that is still several phases
    // behind would fall through (parity 1 of a barrier in phase 0 reads as "the preceding phase", which counts as complete), so the
    // end of the accumulation is signalled on a barrier of its own, used exactly once
    if (gi > 0) {
      if (tid == 0) mma_commit(&full[0]);
      mbar_wait(&full[0], 0);
    }
  } else if (gi > 0) {
    mbar_wait(&bars[(gi - 1) & (S - 1)], ((gi - 1) / S) & 1);
  }
  tc_fence_after();
  // epilogue: thread = row, 16 columns at a time
  const int64_t j = row0 + tid;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int cb = 0; cb < n_tile; cb += 16) {
    uint32_t r[16];
    if (n_it > 0) {
      tmem_ld16(lane_base + cb, r);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = 0;
    }
    if (ksplit > 1) {
      if (j < n_out) {   // this split's partial tile goes to its own slice of the scratch with plain, full-sector stores (no atomics:
                         // the sum over splits is formed in a fixed order by conv_split_finish_kernel)
        uint4* dst = reinterpret_cast<uint4*>(acc + ((int64_t)blockIdx.z * n_out + j) * c_out + n0 + cb);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
      }
    } else if (j < n_out) {
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
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// out[j, c] = bias[c] + sum_z acc[z, j, c]   (offset-split path; fixed summation order)
template <typename T>
__global__ void __launch_bounds__(256)
conv_split_finish_kernel(const float* __restrict__ acc, int n_split, const T* __restrict__ bias, int64_t n_out, int c_out, T* __restrict__ out) {
  const int64_t total = n_out * c_out / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(acc)[i];
    for (int z = 1; z < n_split; ++z) {
      const float4 w = reinterpret_cast<const float4*>(acc)[(int64_t)z * total + i];
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    const int c0 = (int)((i * 4) % c_out);
    float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
    if (bias) { b0 = to_f32(bias[c0]); b1 = to_f32(bias[c0 + 1]); b2 = to_f32(bias[c0 + 2]); b3 = to_f32(bias[c0 + 3]); }
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack2<T>(v.x + b0, v.y + b1), pack2<T>(v.z + b2, v.w + b3));
  }
}

// wt[ci, k, co] = w[co, k, ci]: the backward-data pass as a forward pass over the transposed weights (K-major TMA tiles)
template <typename T>
__global__ void __launch_bounds__(256)
conv_transpose_w_kernel(const T* __restrict__ w, int c_out, int kv, int c_in, T* __restrict__ wt) {
  __shared__ T tile[32][33];
  const int k = blockIdx.z, ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + tx;
    if (co < c_out && ci < c_in) tile[r][tx] = w[((int64_t)co * kv + k) * c_in + ci];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + tx;
    if (co < c_out && ci < c_in) wt[((int64_t)ci * kv + k) * c_out + co] = tile[tx][r];
  }
}

inline bool conv_tma_weights_enabled() {
  static const bool v = [] { const char* e = getenv("B2PC_CONV_TMAW"); return !(e && atoi(e) == 0); }();
  return v;
}
// shapes with compile-time tiles whose weight rows are one swizzle span (kc = 32 -> 64 B, kc = 64 -> 128 B)
inline bool conv_tma_weights_shape(const ConvUmmaCfg& c) {
  return conv_tma_weights_enabled() && (c.kc == 32 || c.kc == 64) &&
         ((c.kc == 32 && c.n_tile == 32) || (c.kc == 64 && (c.n_tile == 64 || c.n_tile == 128 || c.n_tile == 256)));
}

inline size_t conv_umma_workspace_bytes(int64_t n_out, int c_in, int c_out, int kv) {
  if (c_in % 16 != 0 || c_out % 16 != 0) return 0;
  const ConvUmmaCfg c = conv_umma_cfg(n_out, c_in, c_out, kv);
  size_t b = c.ksplit > 1 ? align_up((size_t)c.ksplit * n_out * c_out * sizeof(float), 256) : 0;
  if (conv_tma_weights_shape(c)) b += align_up((size_t)kv * c_in * c_out * 2, 256);   // transposed weights of a backward-data call
  return b ? b + 256 : 0;
}

template <typename T>
inline int launch_gather_gemm_umma_t(const void* feat, const void* weight, const void* bias, const int32_t* pair,
                                     int64_t pair_stride, int64_t n_in, int64_t n_out, int c_in, int c_out, int kv, int transpose_w, int flip,
                                     void* out, void* ws, cudaStream_t stream) {
  const ConvUmmaCfg c = conv_umma_cfg(n_out, c_in, c_out, kv);
  dim3 grid((unsigned)ceil_div(n_out, kCuM), c_out / c.n_tile, c.ksplit);
  float* acc = (float*)ws;
  const bool tb = conv_tma_weights_shape(c);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  int smem_bytes = c.smem_bytes;
  if (tb) {
    const void* wsrc = weight;
    if (transpose_w) {   // weight is [c_in(arg) rows = conv c_out][kv][c_out(arg)]: make the [c_out(arg)][kv][c_in(arg)] copy this pass reads
      T* wt = (T*)((char*)ws + (c.ksplit > 1 ? align_up((size_t)c.ksplit * n_out * c_out * sizeof(float), 256) : 0));
      dim3 tg((unsigned)ceil_div(c_out, 32), (unsigned)ceil_div(c_in, 32), (unsigned)kv);
      conv_transpose_w_kernel<T><<<tg, 256, 0, stream>>>((const T*)weight, c_in, kv, c_out, wt);
      count_launches(1);
      wsrc = wt;
      transpose_w = 0;
    }
    if (!make_tile_tensor_map(&tmap, wsrc, std::is_same<T, __nv_bfloat16>::value, (uint64_t)c_out, (uint64_t)kv * c_in, (uint32_t)c.kc,
                              (uint32_t)c.n_tile, c.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) {
      set_error("spconv_gather_gemm: cuTensorMapEncodeTiled failed");
      return B2PC_ERR_CUDA;
    }
    smem_bytes += 1024;   // alignment slack of the swizzled tiles
  }
  static const int conv_ta = [] { const char* e = getenv("B2PC_CONV_TMAA"); return e ? atoi(e) : 0; }();        // 1: gathered rows by tile::gather4 (correct, but measured slower than cp.async on B200)
  const bool ta = tb && conv_ta != 0 && c_in / c.kc <= kCuMaxChunks;
  ConvGatherMaps tmapf;
  memset(&tmapf, 0, sizeof(tmapf));
  if (ta) {
    for (int i = 0; i < c_in / c.kc; ++i) {
      if (!make_gather4_tensor_map(&tmapf.m[i], (const T*)feat + (size_t)i * c.kc, std::is_same<T, __nv_bfloat16>::value, (uint64_t)n_in,
                                   (uint64_t)c_in, (uint32_t)c.kc, c.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) {
        set_error("spconv_gather_gemm: cuTensorMapEncodeTiled (gather4 map) failed");
        return B2PC_ERR_CUDA;
      }
    }
  }
  static const bool conv_mb = [] { const char* e = getenv("B2PC_CONV_MB"); return e ? atoi(e) != 0 : false; }();   // measured on B200: the block barrier wins (profiles/README.md)
#define B2PC_CONV_LAUNCH_(KC, NT, NCC, MB, TB, TA)                                                                                    \
  do {                                                                                                                             \
    cudaFuncSetAttribute(gather_gemm_umma_kernel<T, KC, NT, NCC, MB, TB, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
    gather_gemm_umma_kernel<T, KC, NT, NCC, MB, TB, TA><<<grid, kCuM, smem_bytes, stream>>>(                                           \
        (const T*)feat, (const T*)weight, (const T*)bias, pair, pair_stride, n_out, c_in, c_out, kv, transpose_w, flip, (T*)out,   \
        c.kc, c.n_tile, c.tmem_cols, c.idx_rows, c.ksplit, acc, tmap, tmapf, (int)n_in);                                                             \
  } while (0)
#define B2PC_CONV_LAUNCH_T(KC, NT, NCC) do { if (ta) B2PC_CONV_LAUNCH_(KC, NT, NCC, false, true, true); else if (tb) B2PC_CONV_LAUNCH_(KC, NT, NCC, false, true, false); else if (conv_mb) B2PC_CONV_LAUNCH_(KC, NT, NCC, true, false, false); else B2PC_CONV_LAUNCH_(KC, NT, NCC, false, false, false); } while (0)
#define B2PC_CONV_LAUNCH(KC, NT, NCC) do { if (conv_mb) B2PC_CONV_LAUNCH_(KC, NT, NCC, true, false, false); else B2PC_CONV_LAUNCH_(KC, NT, NCC, false, false, false); } while (0)
  const int ncc = c_in / c.kc;
  if (c.kc == 32 && c.n_tile == 32 && ncc == 1) B2PC_CONV_LAUNCH_T(32, 32, 1);
  else if (c.kc == 64 && c.n_tile == 64 && ncc == 1) B2PC_CONV_LAUNCH_T(64, 64, 1);
  else if (c.kc == 64 && c.n_tile == 128 && ncc == 2) B2PC_CONV_LAUNCH_T(64, 128, 2);
  else if (c.kc == 64 && c.n_tile == 256 && ncc == 4) B2PC_CONV_LAUNCH_T(64, 256, 4);
  else if (c.kc == 64 && c.n_tile == 64) B2PC_CONV_LAUNCH_T(64, 64, 0);
  else if (c.kc == 64 && c.n_tile == 128) B2PC_CONV_LAUNCH_T(64, 128, 0);
  else if (c.kc == 64 && c.n_tile == 256) B2PC_CONV_LAUNCH_T(64, 256, 0);
  else if (c.kc == 16 && c.n_tile == 32 && ncc == 1) B2PC_CONV_LAUNCH(16, 32, 1);
  else B2PC_CONV_LAUNCH(0, 0, 0);
#undef B2PC_CONV_LAUNCH_T
#undef B2PC_CONV_LAUNCH
#undef B2PC_CONV_LAUNCH_
  if (c.ksplit > 1) {
    int64_t fb = ceil_div(n_out * c_out / 4, 256);
    if (fb > kNumSMs * 8) fb = kNumSMs * 8;
    conv_split_finish_kernel<T><<<(int)fb, 256, 0, stream>>>(acc, c.ksplit, (const T*)bias, n_out, c_out, (T*)out);
    count_launches(1);
  }
  count_launches(1);
  B2PC_CHECK_LAUNCH("spconv_gather_gemm(tcgen05)");
  return B2PC_OK;
}

inline int launch_gather_gemm_umma(const void* feat, const void* weight, const void* bias, const int32_t* pair, int64_t pair_stride,
                                   int64_t n_in, int64_t n_out, int c_in, int c_out, int kv, int transpose_w, int flip, int dtype,
                                   void* out, void* ws, cudaStream_t stream) {
  if (n_out == 0) return B2PC_OK;
  if (dtype == B2PC_BF16)
    return launch_gather_gemm_umma_t<__nv_bfloat16>(feat, weight, bias, pair, pair_stride, n_in, n_out, c_in, c_out, kv, transpose_w, flip, out, ws, stream);
  return launch_gather_gemm_umma_t<__half>(feat, weight, bias, pair, pair_stride, n_in, n_out, c_in, c_out, kv, transpose_w, flip, out, ws, stream);
}

}  // namespace b2pc

namespace b2pc {

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:   dW[co, k, ci] = sum_j dout[j, co] * feat[pair[k, j], ci]
// A CTA owns a group of kernel offsets (as many [M x N] fp32 accumulators as fit in the 512 TMEM columns), an M tile of
// output channels and an N tile of input channels, and sweeps its share of 64-row tiles of the rulebook.  Per row tile the
// dout slice is staged once (operand A, MN-major: channels contiguous, rows are the reduction axis) and for every offset
// with at least one partner in the tile the gathered feature rows are staged as operand B (MN-major as well); 4
// tcgen05.mma (K = 16 rows each) accumulate into that offset's TMEM slice.  Row-range splits are reduced afterwards in a
// fixed order (deterministic, no atomics).
constexpr int kWuRows = 64;      // rulebook rows per step (reduction chunk)
constexpr int kWuStages = 4;

struct WgradCfg { int m_tile, n_mtiles, n_tile, n_ntiles, g_size, n_groups, n_splits, smem_bytes, tmem_cols; };

inline WgradCfg wgrad_cfg(int64_t n_out, int c_in, int c_out, int kv) {
  WgradCfg c;
  c.m_tile = c_out <= 64 ? 64 : 128;
  c.n_mtiles = (c_out + c.m_tile - 1) / c.m_tile;
  c.n_tile = c_in <= 256 ? c_in : (c_in % 256 == 0 ? 256 : (c_in % 128 == 0 ? 128 : 64));
  c.n_ntiles = c_in / c.n_tile;
  int ncol = 32;
  while (ncol < c.n_tile) ncol <<= 1;          // keep accumulator slices on power-of-two column strides
  // TMEM budget per CTA: narrow tiles take 128 columns so that four CTAs share an SM (their per-step work is tiny and
  // latency bound), wide tiles take more columns and rely on the tensor pipe instead
  c.tmem_cols = ncol <= 64 ? 128 : (ncol == 128 ? 256 : 512);   // (fewer, larger groups for 5^3 kernels measured 2.3x slower)
  c.g_size = c.tmem_cols / ncol;
  if (c.g_size > kv) c.g_size = kv;
  if (c.g_size > 32) c.g_size = 32;
  c.n_groups = (kv + c.g_size - 1) / c.g_size;
  const int64_t tiles = ceil_div(n_out, kWuRows);
  const int per_sm = 512 / c.tmem_cols;
  int64_t sp = ceil_div((int64_t)2 * per_sm * kNumSMs, (int64_t)c.n_groups * c.n_mtiles * c.n_ntiles);
  if (sp > tiles) sp = tiles;
  if (sp < 1) sp = 1;
  c.n_splits = (int)sp;
  c.smem_bytes = 32 * kWuRows * 4 + 256 + kWuRows * c.m_tile * 2 + kWuStages * kWuRows * c.n_tile * 2;
  return c;
}

inline bool wgrad_umma_supported(int dtype, int c_in, int c_out) {
  if (dtype != B2PC_F16 && dtype != B2PC_BF16) return false;
  if (c_in % 16 != 0 || c_out % 8 != 0) return false;
  if (c_in > 256 && c_in % 64 != 0) return false;
  return true;
}

inline size_t wgrad_umma_workspace_bytes(int64_t n_out, int c_in, int c_out, int kv) {
  if (c_in % 16 != 0 || c_out % 8 != 0 || (c_in > 256 && c_in % 64 != 0)) return 0;
  const WgradCfg c = wgrad_cfg(n_out > 0 ? n_out : 1, c_in, c_out, kv);
  return (size_t)c.n_splits * c_out * kv * c_in * sizeof(float) + 256;
}

template <typename T>
__global__ void __launch_bounds__(128)
bwd_weight_umma_kernel(const T* __restrict__ feat, const T* __restrict__ dout, const int32_t* __restrict__ pair,
                       int64_t pair_stride, int64_t n_out, int c_in, int c_out, int kv, float* __restrict__ partial, WgradCfg cfg) {
  using namespace umma;
  extern __shared__ __align__(128) uint8_t smem[];
  int32_t* idx_s = reinterpret_cast<int32_t*>(smem);                        // [32][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32 * kWuRows * 4);    // [kWuStages] stage free, [kWuStages] = tile done
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(bars + kWuStages + 1);
  uint32_t* tmem_slot = mask_s + 1;
  uint8_t* a_s = smem + 32 * kWuRows * 4 + 256;
  const int a_bytes = kWuRows * cfg.m_tile * 2, b_bytes = kWuRows * cfg.n_tile * 2;
  uint8_t* b_s0 = a_s + a_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int bx = blockIdx.x;
  const int nt = bx % cfg.n_ntiles; bx /= cfg.n_ntiles;
  const int mt = bx % cfg.n_mtiles; bx /= cfg.n_mtiles;
  const int group = bx;
  const int split = blockIdx.y;
  const int k_begin = group * cfg.g_size;
  const int k_cnt = min(cfg.g_size, kv - k_begin);
  const int co0 = mt * cfg.m_tile, ci0 = nt * cfg.n_tile;
  const int m_valid = min(cfg.m_tile, c_out - co0);
  int ncol = 32;
  while (ncol < cfg.n_tile) ncol <<= 1;

  if (warp == 0) { tmem_alloc(tmem_slot, cfg.tmem_cols); tmem_relinquish(); }
  if (tid == 0) {
    for (int s = 0; s <= kWuStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  // zero the A planes that no dout channel will ever fill (c_out smaller than the M tile)
  for (int q = tid; q < a_bytes / 16; q += 128) reinterpret_cast<uint4*>(a_s)[q] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc(cfg.m_tile, cfg.n_tile, UmmaFmt<T>::v, UmmaFmt<T>::v, 1, 1);

  const int64_t n_tiles = ceil_div(n_out, kWuRows);
  uint32_t inited = 0;        // offsets (bit = k - k_begin) whose accumulator holds data
  uint32_t n_commit[kWuStages] = {0, 0, 0, 0};   // completed uses per B stage (uniform across threads)
  uint32_t n_tile_commit = 0;
  int ring = 0;               // next B stage

  for (int64_t t = split; t < n_tiles; t += cfg.n_splits) {
    const int64_t r0 = t * kWuRows;
    // (1) previous tile's MMAs must be done before A / idx are overwritten
    if (n_tile_commit > 0) mbar_wait(&bars[kWuStages], (n_tile_commit - 1) & 1);
    if (tid == 0) *mask_s = 0;
    __syncthreads();
    // (2) rulebook slice + active-offset mask
    for (int q = tid; q < k_cnt * kWuRows; q += 128) {
      const int kk = q / kWuRows, r = q % kWuRows;
      const int64_t j = r0 + r;
      const int32_t v = j < n_out ? pair[(int64_t)(k_begin + kk) * pair_stride + j] : -1;
      idx_s[kk * kWuRows + r] = v;
      const unsigned b = __ballot_sync(0xFFFFFFFFu, v >= 0);   // a warp covers rows of one offset (64 rows = 2 warps)
      if (lane == 0 && b) atomicOr(mask_s, 1u << kk);
    }
    // (3) A = dout rows of this tile: piece (row, chunk p) -> p*1024 + row*16
    {
      const int ppr = m_valid / 8;
      for (int q = tid; q < kWuRows * ppr; q += 128) {
        const int r = q / ppr, p = q % ppr;
        const int64_t j = r0 + r;
        cp_async16(smem_u32(a_s) + p * (kWuRows * 16) + r * 16, dout + (j < n_out ? j : 0) * c_out + co0 + p * 8, j < n_out);
      }
      cp_async_commit();
    }
    __syncthreads();
    const uint32_t mask = *mask_s;
    const int n_act = __popc(mask);
    // (4) sweep the active offsets with a small ring of gathered-B stages
    auto gather = [&](int a_i, int stage) {
      const int kk = __fns(mask, 0, a_i + 1);
      const int r = tid >> 1, half = tid & 1;
      const int32_t src = idx_s[kk * kWuRows + r];
      const T* g = feat + (int64_t)(src >= 0 ? src : 0) * c_in + ci0;
      const uint32_t dst = smem_u32(b_s0 + stage * b_bytes) + r * 16;
      for (int p = half; p < cfg.n_tile / 8; p += 2) cp_async16(dst + p * (kWuRows * 16), g + p * 8, src >= 0);
    };
    // prologue: up to 2 gathers in flight
    for (int a_i = 0; a_i < 2; ++a_i) {
      if (a_i < n_act) {
        const int s = (ring + a_i) % kWuStages;
        if (n_commit[s] > 0) mbar_wait(&bars[s], (n_commit[s] - 1) & 1);
        gather(a_i, s);
      }
      cp_async_commit();
    }
    for (int a_i = 0; a_i < n_act; ++a_i) {
      const int s = (ring + a_i) % kWuStages;
      if (a_i + 2 < n_act) {
        const int s2 = (ring + a_i + 2) % kWuStages;
        if (n_commit[s2] > 0) mbar_wait(&bars[s2], (n_commit[s2] - 1) & 1);
        gather(a_i + 2, s2);
      }
      cp_async_commit();
      cp_async_wait<2>();     // A tile + gather a_i have landed (groups: A, g0, g1, then one per iteration)
      fence_proxy_async();
      __syncthreads();
      const int kk = __fns(mask, 0, a_i + 1);
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s0 + s * b_bytes);
#pragma unroll
        for (int ks = 0; ks < kWuRows / 16; ++ks)
          mma_ss(tmem_base + kk * ncol, make_smem_desc(a_addr + ks * 256, 128, kWuRows * 16),
                 make_smem_desc(b_addr + ks * 256, 128, kWuRows * 16), idesc, (((inited >> kk) & 1u) || ks > 0) ? 1u : 0u);
        mma_commit(&bars[s]);
        if (a_i == n_act - 1) mma_commit(&bars[kWuStages]);
      }
      inited |= 1u << kk;
      n_commit[s] += 1;
    }
    if (n_act > 0) n_tile_commit += 1;
    ring = (ring + n_act) % kWuStages;
    cp_async_wait<0>();
  }
  if (n_tile_commit > 0) mbar_wait(&bars[kWuStages], (n_tile_commit - 1) & 1);
  tc_fence_after();
  // epilogue: accumulator row -> partial[split][co][k][ci0 .. ci0+n_tile)
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  int row;
  bool row_ok;
  if (cfg.m_tile == 128) { row = tid; row_ok = row < m_valid; }
  else { row = warp * 16 + lane; row_ok = lane < 16 && row < m_valid; }
  for (int kk = 0; kk < k_cnt; ++kk) {
    const bool has = (inited >> kk) & 1u;
    float* dst = partial + (((int64_t)split * c_out + co0 + row) * kv + k_begin + kk) * c_in + ci0;
    for (int cb = 0; cb < cfg.n_tile; cb += 16) {
      uint32_t r[16];
      if (has) {
        tmem_ld16(lane_base + kk * ncol + cb, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0;
      }
      if (row_ok) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          reinterpret_cast<float4*>(dst + cb)[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]),
                                                                __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, cfg.tmem_cols);
}

template <typename T>
inline int launch_bwd_weight_umma_t(const void* feat, const void* dout, const int32_t* pair, int64_t pair_stride, int64_t n_out,
                                    int c_in, int c_out, int kv, float* dweight, void* ws, cudaStream_t stream) {
  const WgradCfg c = wgrad_cfg(n_out, c_in, c_out, kv);
  cudaFuncSetAttribute(bwd_weight_umma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes);   // per device, so per launch
  dim3 grid(c.n_groups * c.n_mtiles * c.n_ntiles, c.n_splits);
  bwd_weight_umma_kernel<T><<<grid, 128, c.smem_bytes, stream>>>((const T*)feat, (const T*)dout, pair, pair_stride, n_out, c_in, c_out,
                                                                 kv, (float*)ws, c);
  const int64_t elems = (int64_t)c_out * kv * c_in;
  int rb = (int)ceil_div(elems, 256);
  if (rb > kNumSMs * 8) rb = kNumSMs * 8;
  reduce_splits_kernel<<<rb, 256, 0, stream>>>((const float*)ws, elems, c.n_splits, dweight);
  count_launches(2);
  B2PC_CHECK_LAUNCH("spconv_bwd_weight(tcgen05)");
  return B2PC_OK;
}

inline int launch_bwd_weight_umma(const void* feat, const void* dout, const int32_t* pair, int64_t pair_stride, int64_t n_out, int c_in,
                                  int c_out, int kv, int dtype, float* dweight, void* ws, cudaStream_t stream) {
  if (dtype == B2PC_BF16)
    return launch_bwd_weight_umma_t<__nv_bfloat16>(feat, dout, pair, pair_stride, n_out, c_in, c_out, kv, dweight, ws, stream);
  return launch_bwd_weight_umma_t<__half>(feat, dout, pair, pair_stride, n_out, c_in, c_out, kv, dweight, ws, stream);
}

}  // namespace b2pc
