// Operators of the variants that share the hot-path kernels:
//   knn_query_kernel       pointops.knn_query (libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104), used to map predictions
//                          back to the original points (pointcept/engines/hooks/evaluator.py:569-575) and by kNN interpolation
//   vote_accumulate_kernel fragment voting of the tester: pred[idx_part] += softmax(seg_logits) (pointcept/engines/test.py:193-203)
//   point_rope_kernel      PointROPE (libs/pointrope/kernels.cu:19-76): in-place 3-axis rotary embedding of q / k with integer voxel
//                          positions (LitePT, pointcept/models/litept/litept_v1.py:231-241; PT-v3m3's RoPE variant)
#pragma once
#include "common.cuh"

namespace b2pc {

// ---- kNN ---------------------------------------------------------------------------------------------------------------------
// Brute force like the reference (exact, any k <= 128), but the candidate points are staged through shared memory once per block of
// 128 queries instead of being re-read from global memory by every query thread, and the running top-k is a sorted list.
constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 1024;

template <int KMAX>
__global__ void __launch_bounds__(kKnnThreads)
knn_query_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offset, const float* __restrict__ new_xyz,
                 const int32_t* __restrict__ new_offset, int nb, int64_t m, int k, int32_t* __restrict__ idx, float* __restrict__ dist2) {
  __shared__ float4 tile[kKnnTile];
  __shared__ int range_s[2];
  const int64_t q = blockIdx.x * (int64_t)kKnnThreads + threadIdx.x;
  const bool ok = q < m;
  int start = 0, end = 0;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (ok) {
    int b = 0;
    while (b < nb - 1 && q >= new_offset[b]) ++b;      // get_bt_idx, knn_query_cuda_kernel.cu:47-57
    start = b ? offset[b - 1] : 0;
    end = offset[b];
    qx = new_xyz[q * 3 + 0]; qy = new_xyz[q * 3 + 1]; qz = new_xyz[q * 3 + 2];
  }
  // candidate range of the block = union over its queries (queries are grouped by scene, so this is one or two scenes)
  if (threadIdx.x == 0) { range_s[0] = INT_MAX; range_s[1] = 0; }
  __syncthreads();
  if (ok) { atomicMin(&range_s[0], start); atomicMax(&range_s[1], end); }
  __syncthreads();
  const int lo = range_s[0], hi = range_s[1];
  float bd[KMAX];
  int bi[KMAX];
#pragma unroll
  for (int i = 0; i < KMAX; ++i) { bd[i] = 1e10f; bi[i] = -1; }
  float worst = 1e10f;
  for (int t0 = lo; t0 < hi; t0 += kKnnTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKnnTile; e += kKnnThreads) {
      const int i = t0 + e;
      tile[e] = i < hi ? make_float4(xyz[(int64_t)i * 3], xyz[(int64_t)i * 3 + 1], xyz[(int64_t)i * 3 + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int e0 = max(start, t0) - t0, e1 = min(end, t0 + kKnnTile) - t0;
    for (int e = e0; e < e1; ++e) {
      const float4 p = tile[e];
      const float d2 = (qx - p.x) * (qx - p.x) + (qy - p.y) * (qy - p.y) + (qz - p.z) * (qz - p.z);
      if (d2 < worst) {          // strictly closer than the current k-th: insert behind equal distances (earlier index first)
        bool placed = false;
#pragma unroll
        for (int s = KMAX - 1; s >= 0; --s) {      // static indices only: the list stays in registers
          if (s < k && !placed) {
            if (s > 0 && bd[s - 1] > d2) { bd[s] = bd[s - 1]; bi[s] = bi[s - 1]; }
            else { bd[s] = d2; bi[s] = t0 + e; placed = true; }
          }
        }
#pragma unroll
        for (int s = 0; s < KMAX; ++s)
          if (s == k - 1) worst = bd[s];
      }
    }
  }
  if (ok) {
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
      if (i < k) { idx[q * k + i] = bi[i]; dist2[q * k + i] = bd[i]; }
  }
}

inline int launch_knn_query(const float* xyz, const int32_t* offset, const float* new_xyz, const int32_t* new_offset, int nb, int64_t m, int k,
                            int32_t* idx, float* dist2, cudaStream_t stream) {
  B2PC_CHECK_ARG(nb >= 1 && m >= 0 && k >= 1 && k <= 128, "knn_query: nsample must be in [1,128] (got %d)", k);
  if (m == 0) return B2PC_OK;
  const unsigned blocks = (unsigned)ceil_div(m, kKnnThreads);
#define B2PC_KNN(KM) knn_query_kernel<KM><<<blocks, kKnnThreads, 0, stream>>>(xyz, offset, new_xyz, new_offset, nb, m, k, idx, dist2)
  if (k == 1) B2PC_KNN(1);
  else if (k <= 4) B2PC_KNN(4);
  else if (k <= 8) B2PC_KNN(8);
  else if (k <= 16) B2PC_KNN(16);
  else if (k <= 32) B2PC_KNN(32);
  else if (k <= 64) B2PC_KNN(64);
  else B2PC_KNN(128);
#undef B2PC_KNN
  count_launches(1);
  B2PC_CHECK_LAUNCH("knn_query");
  return B2PC_OK;
}

// ---- fragment voting -----------------------------------------------------------------------------------------------------------
// pred[index[i], :] += softmax(logits[i, :]); one warp per row, classes strided over the lanes (fp32 softmax like F.softmax on fp32)
template <typename T>
__global__ void __launch_bounds__(256)
vote_accumulate_kernel(const T* __restrict__ logits, const int64_t* __restrict__ index, int64_t n, int c, float* __restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const T* x = logits + row * c;
  float mx = -INFINITY;
  for (int j = lane; j < c; j += 32) mx = fmaxf(mx, to_f32(x[j]));
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < c; j += 32) sum += expf(to_f32(x[j]) - mx);
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
  const float inv = 1.f / sum;
  float* dst = pred + index[row] * c;
  for (int j = lane; j < c; j += 32) atomicAdd(dst + j, expf(to_f32(x[j]) - mx) * inv);
}

// ---- PointROPE -------------------------------------------------------------------------------------------------------------------
// tokens: n_tok tokens of n_heads heads of D channels (token stride given, heads contiguous), rotated in place.  Channel layout of a
// head (kernels.cu:40-42), Q = D/6:  [u_X(Q) v_X(Q) u_Y(Q) v_Y(Q) u_Z(Q) v_Z(Q)];  angle = pos[axis] * fwd / base^(q/Q);
//   u' = u cos - v sin,  v' = v cos + u sin.   One thread per (token, axis, q): the sin/cos pair is shared by all heads.
template <typename T>
__global__ void __launch_bounds__(256)
point_rope_kernel(T* __restrict__ tokens, const int64_t* __restrict__ pos, int64_t n_tok, int64_t token_stride, int n_heads, int d,
                  float base, float fwd) {
  const int q_per = d / 6, per_tok = 3 * q_per;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_tok * per_tok) return;
  const int64_t t = e / per_tok;
  const int r = (int)(e - t * per_tok), axis = r / q_per, q = r - axis * q_per;
  const float inv_freq = fwd / powf(base, q / float(q_per));
  const float ang = (float)pos[t * 3 + axis] * inv_freq;
  const float c = cosf(ang), s = sinf(ang);
  T* p = tokens + t * token_stride + axis * (d / 3) + q;
  for (int h = 0; h < n_heads; ++h, p += d) {
    const float u = to_f32(p[0]), v = to_f32(p[q_per]);
    p[0] = from_f32<T>(u * c - v * s);
    p[q_per] = from_f32<T>(v * c + u * s);
  }
}

}  // namespace b2pc
