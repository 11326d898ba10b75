// Space-filling-curve encoding of voxel coordinates: all requested orders in one pass.
// HBM-bound: 12 B (xyz) + 8 B (batch) in, 8 B per order out, per point.
#pragma once
#include "common.cuh"

namespace b2pc {

// Spread the low 16 bits of v so that bit i lands on bit 3*i (magic-number Morton spread, 64-bit).
__device__ __forceinline__ uint64_t spread3(uint32_t v) {
  uint64_t x = v & 0xFFFFull;
  x = (x | (x << 16)) & 0x0000FF0000FFull;      // never crosses: 16 bits -> two bytes 32.. apart
  x = (x | (x << 8)) & 0x00F00F00F00Full;
  x = (x | (x << 4)) & 0x0C30C30C30C3ull;
  x = (x | (x << 2)) & 0x249249249249ull;
  return x;
}

// Morton key with x in the most significant bit of each triple (z_order.py:39-49).
__device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
  return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}

// Skilling's axes->transpose on integer lanes followed by interleave + Gray->binary prefix xor.
// Semantics of hilbert.py:150-190: bits MSB first, lanes 0,1,2 in order; "bit set" inverts the
// lower bits of lane 0, "bit clear" exchanges the differing lower bits of lane 0 and lane d.
__device__ __forceinline__ uint64_t hilbert3(uint32_t x0, uint32_t x1, uint32_t x2, int depth) {
  uint32_t X[3] = {x0, x1, x2};
#pragma unroll 1
  for (uint32_t q = 1u << (depth - 1); q > 0; q >>= 1) {
    const uint32_t p = q - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (X[d] & q) {
        X[0] ^= p;
      } else {
        const uint32_t t = (X[0] ^ X[d]) & p;
        X[0] ^= t;
        X[d] ^= t;
      }
    }
  }
  uint64_t g = morton3(X[0], X[1], X[2]);
  g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8; g ^= g >> 16; g ^= g >> 32;
  return g;
}

struct EncodeOrders { int id[8]; int n; };

__global__ void __launch_bounds__(256)
encode_kernel(const int32_t* __restrict__ grid, const int64_t* __restrict__ batch, int64_t n, int depth,
              EncodeOrders orders, int64_t* __restrict__ code) {
  const uint32_t mask = (depth >= 32) ? 0xFFFFFFFFu : ((1u << depth) - 1u);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)grid[3 * i + 0] & mask;
    const uint32_t y = (uint32_t)grid[3 * i + 1] & mask;
    const uint32_t z = (uint32_t)grid[3 * i + 2] & mask;
    const uint64_t hi = batch ? ((uint64_t)batch[i] << (3 * depth)) : 0ull;
#pragma unroll 1
    for (int o = 0; o < orders.n; ++o) {
      uint64_t c;
      switch (orders.id[o]) {
        case B2PC_ORDER_Z: c = morton3(x, y, z); break;
        case B2PC_ORDER_Z_TRANS: c = morton3(y, x, z); break;
        case B2PC_ORDER_HILBERT: c = hilbert3(x, y, z, depth); break;
        default: c = hilbert3(y, x, z, depth); break;
      }
      code[(int64_t)o * n + i] = (int64_t)(hi | c);
    }
  }
}

inline int launch_encode(const int32_t* grid, const int64_t* batch, int64_t n, int depth, const int* orders,
                         int n_orders, int64_t* code, cudaStream_t stream) {
  B2PC_CHECK_ARG(depth >= 1 && depth <= 16, "serialize_encode: depth %d outside [1,16]", depth);
  B2PC_CHECK_ARG(n_orders >= 1 && n_orders <= 8, "serialize_encode: n_orders %d outside [1,8]", n_orders);
  B2PC_CHECK_ARG(n >= 0, "serialize_encode: negative n");
  EncodeOrders eo;
  eo.n = n_orders;
  for (int o = 0; o < n_orders; ++o) {
    B2PC_CHECK_ARG(orders[o] >= 0 && orders[o] <= 3, "serialize_encode: unknown order id %d", orders[o]);
    eo.id[o] = orders[o];
  }
  if (n == 0) return B2PC_OK;
  int blocks = (int)ceil_div(n, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  encode_kernel<<<blocks, 256, 0, stream>>>(grid, batch, n, depth, eo, code);
  count_launches(1);
  B2PC_CHECK_LAUNCH("serialize_encode");
  return B2PC_OK;
}

// ---- patch padding tables (ptv3m1:114-170) -----------------------------------------------------
// One thread per padded slot / per point; scene found by binary search in the (tiny) offset table.
__global__ void __launch_bounds__(256)
padding_kernel(const int64_t* __restrict__ offset, int B, int K, int64_t n, int64_t t_pad, int n_seq,
               int64_t* __restrict__ pad, int64_t* __restrict__ unpad, int32_t* __restrict__ cu) {
  extern __shared__ int64_t sh[];  // [0..B] raw starts, [B+1 .. 2B+1] padded starts
  int64_t* o = sh;
  int64_t* op = sh + (B + 1);
  if (threadIdx.x == 0) {
    o[0] = 0; op[0] = 0;
    for (int b = 0; b < B; ++b) {
      o[b + 1] = offset[b];
      const int64_t c = o[b + 1] - o[b];
      op[b + 1] = op[b] + (c > K ? (c + K - 1) / K * K : c);
    }
  }
  __syncthreads();
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = gid; t < t_pad; t += stride) {
    int lo = 0, hi = B;  // largest b with op[b] <= t
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (op[mid] <= t) lo = mid; else hi = mid; }
    const int b = lo;
    const int64_t cnt = o[b + 1] - o[b];
    const int64_t cntp = op[b + 1] - op[b];
    int64_t local = t - op[b];
    if (cnt != cntp) {
      const int64_t r = cnt % K;
      if (local >= cntp - K + r) local -= K;  // borrowed tail: copy of the slot K earlier
    }
    pad[t] = o[b] + local;
  }
  for (int64_t i = gid; i < n; i += stride) {
    int lo = 0, hi = B;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (o[mid] <= i) lo = mid; else hi = mid; }
    unpad[i] = i + (op[lo] - o[lo]);
  }
  // cu_seqlens: one entry per patch start, plus the total
  for (int64_t s = gid; s <= n_seq; s += stride) {
    if (s == n_seq) { cu[s] = (int32_t)t_pad; continue; }
    // patch s -> scene: count patches per scene
    int64_t acc = 0; int b = 0;
    for (; b < B; ++b) {
      const int64_t cntp = op[b + 1] - op[b];
      const int64_t np = cntp == 0 ? 0 : (cntp + K - 1) / K;
      if (s < acc + np) break;
      acc += np;
    }
    cu[s] = (int32_t)(op[b] + (s - acc) * K);
  }
}

inline int launch_padding(const int64_t* offset, int B, int K, int64_t n, int64_t t_pad, int n_seq, int64_t* pad,
                          int64_t* unpad, int32_t* cu, cudaStream_t stream) {
  B2PC_CHECK_ARG(B >= 1 && B <= 2048, "patch_padding: batch_size %d outside [1,2048]", B);
  B2PC_CHECK_ARG(K >= 1, "patch_padding: patch_size must be positive");
  B2PC_CHECK_ARG(t_pad >= n && t_pad < (1ll << 31), "patch_padding: t_pad %lld invalid", (long long)t_pad);
  int64_t work = t_pad > n ? t_pad : n;
  int blocks = (int)ceil_div(work > 0 ? work : 1, 256);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  padding_kernel<<<blocks, 256, (2 * B + 2) * sizeof(int64_t), stream>>>(offset, B, K, n, t_pad, n_seq, pad, unpad, cu);
  count_launches(1);
  B2PC_CHECK_LAUNCH("patch_padding");
  return B2PC_OK;
}

}  // namespace b2pc
