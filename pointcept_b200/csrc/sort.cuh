// Segmented stable LSD radix sort of (int64 code -> int32 row) pairs, 8-bit digits, only over the
// significant key bits.  One grid row (blockIdx.y) per serialization order, so all orders are sorted
// by the same launches.  Per pass and key: 8 B read (histogram) + 12 B read + 12 B write.
#pragma once
#include "common.cuh"

namespace b2pc {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                       // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per block
constexpr int kSortWarps = kSortThreads / 32;

// counts[seg][digit][block]
__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int nblocks, uint32_t* __restrict__ counts) {
  __shared__ uint32_t h[256];
  const int seg = blockIdx.y;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t* k = keys + (int64_t)seg * n;
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t i = base + it * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(k[i] >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  counts[((int64_t)seg * 256 + threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan over (digit major, block minor) for each segment; one block per segment.
__global__ void __launch_bounds__(1024)
sort_scan_kernel(uint32_t* __restrict__ counts, int nblocks) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  uint32_t* c = counts + (int64_t)blockIdx.x * 256 * nblocks;
  const int64_t total = (int64_t)256 * nblocks;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < total; base += 1024) {
    const int64_t i = base + threadIdx.x;
    uint32_t v = (i < total) ? c[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u;
    const uint32_t carry = carry_s;
    if (i < total) c[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
}

// Stable scatter.  Element order inside a tile is (warp, round, lane): warp w owns the contiguous
// slice [w*256, (w+1)*256) of the tile and walks it 32 keys per round.
template <bool kFirst, bool kLast>
__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                    int64_t* __restrict__ order_out, int64_t* __restrict__ inverse_out,
                    int64_t n, int shift, int nblocks, const uint32_t* __restrict__ offsets) {
  __shared__ uint32_t warp_digit[kSortWarps][256];  // per-warp running count per digit, then warp base
  __shared__ uint32_t block_base[256];
  const int seg = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&warp_digit[0][0])[i] = 0;
  block_base[threadIdx.x] = offsets[((int64_t)seg * 256 + threadIdx.x) * nblocks + blockIdx.x];
  __syncthreads();

  const uint64_t* kin = keys_in + (int64_t)seg * n;
  const int64_t tile0 = (int64_t)blockIdx.x * kSortTile + warp * (32 * kSortItems);
  uint64_t key[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t i = tile0 + it * 32 + lane;
    const bool ok = i < n;
    key[it] = ok ? kin[i] : ~0ull;
    const uint32_t d = (uint32_t)(key[it] >> shift) & 0xFF;
    // peers = lanes of this round holding the same digit (invalid lanes grouped under digit 256)
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ok ? d : 256u);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    uint32_t prev = 0;
    if (ok && lane == leader) {
      prev = warp_digit[warp][d];
      warp_digit[warp][d] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
    rank[it] = prev + before;
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over warps per digit -> warp base inside the block's digit bucket
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = warp_digit[w][d];
      warp_digit[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  const int64_t seg_base = (int64_t)seg * n;
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t i = tile0 + it * 32 + lane;
    if (i < n) {
      const uint32_t d = (uint32_t)(key[it] >> shift) & 0xFF;
      const int64_t pos = (int64_t)block_base[d] + warp_digit[warp][d] + rank[it];
      const uint32_t v = kFirst ? (uint32_t)i : vals_in[seg_base + i];
      if (kLast) {
        order_out[seg_base + pos] = (int64_t)v;
        inverse_out[seg_base + v] = pos;
      } else {
        keys_out[seg_base + pos] = key[it];
        vals_out[seg_base + pos] = v;
      }
    }
  }
}

inline size_t sort_workspace_bytes(int64_t n, int n_orders) {
  const int64_t nblocks = ceil_div(n > 0 ? n : 1, kSortTile);
  size_t b = 0;
  b += align_up((size_t)n_orders * n * 8, 256) * 2;      // key ping-pong
  b += align_up((size_t)n_orders * n * 4, 256) * 2;      // value ping-pong
  b += align_up((size_t)n_orders * 256 * nblocks * 4, 256);
  return b;
}

inline int launch_sort(const int64_t* code, int64_t n, int n_orders, int key_bits, int64_t* order, int64_t* inverse,
                       void* ws, size_t ws_bytes, cudaStream_t stream) {
  B2PC_CHECK_ARG(n >= 0 && n < (1ll << 31), "serialize_sort: n %lld out of range", (long long)n);
  B2PC_CHECK_ARG(n_orders >= 1 && n_orders <= 4096, "serialize_sort: n_orders %d outside [1,4096]", n_orders);
  B2PC_CHECK_ARG(key_bits >= 1 && key_bits <= 64, "serialize_sort: key_bits %d outside [1,64]", key_bits);
  if (ws_bytes < sort_workspace_bytes(n, n_orders)) {
    set_error("serialize_sort: workspace %zu < required %zu", ws_bytes, sort_workspace_bytes(n, n_orders));
    return B2PC_ERR_WORKSPACE;
  }
  if (n == 0) return B2PC_OK;
  const int nblocks = (int)ceil_div(n, kSortTile);
  char* p = (char*)ws;
  uint64_t* kbuf[2]; uint32_t* vbuf[2];
  kbuf[0] = (uint64_t*)p; p += align_up((size_t)n_orders * n * 8, 256);
  kbuf[1] = (uint64_t*)p; p += align_up((size_t)n_orders * n * 8, 256);
  vbuf[0] = (uint32_t*)p; p += align_up((size_t)n_orders * n * 4, 256);
  vbuf[1] = (uint32_t*)p; p += align_up((size_t)n_orders * n * 4, 256);
  uint32_t* counts = (uint32_t*)p;
  const int passes = (key_bits + 7) / 8;
  const uint64_t* kin = (const uint64_t*)code;
  const uint32_t* vin = nullptr;
  dim3 grid(nblocks, n_orders);
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * 8;
    const bool first = pass == 0, last = pass == passes - 1;
    sort_hist_kernel<<<grid, kSortThreads, 0, stream>>>(kin, n, shift, nblocks, counts);
    sort_scan_kernel<<<n_orders, 1024, 0, stream>>>(counts, nblocks);
    uint64_t* kout = kbuf[pass & 1];
    uint32_t* vout = vbuf[pass & 1];
    if (first && last)
      sort_scatter_kernel<true, true><<<grid, kSortThreads, 0, stream>>>(kin, vin, kout, vout, order, inverse, n, shift, nblocks, counts);
    else if (first)
      sort_scatter_kernel<true, false><<<grid, kSortThreads, 0, stream>>>(kin, vin, kout, vout, order, inverse, n, shift, nblocks, counts);
    else if (last)
      sort_scatter_kernel<false, true><<<grid, kSortThreads, 0, stream>>>(kin, vin, kout, vout, order, inverse, n, shift, nblocks, counts);
    else
      sort_scatter_kernel<false, false><<<grid, kSortThreads, 0, stream>>>(kin, vin, kout, vout, order, inverse, n, shift, nblocks, counts);
    kin = kout;
    vin = vout;
    count_launches(3);
  }
  B2PC_CHECK_LAUNCH("serialize_sort");
  return B2PC_OK;
}

}  // namespace b2pc
