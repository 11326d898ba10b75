// Micro-probe of cp.async.bulk.tensor tile::gather4 semantics on sm_100a (run under gpurun): what lands in shared memory for
//  (A) a map over the full [R x 128] matrix, box {64,1}, column coordinate 0 / 64;
//  (B) a map over a 64-column window (base advanced by 0 / 64 elements, row stride 256 B), column coordinate 0.
// Element (r, c) holds r * 256 + c, so every 16-bit value identifies its source.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int col, int r0, int r1, int r2, int r3, uint16_t* out) {
  __shared__ __align__(1024) uint16_t tile[4 * 64];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tile[i] = 0xFFFF;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4 * 64 * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(tile)), "l"(&map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    for (int polls = 0; !done && polls < (1 << 22); ++polls)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    if (!done) tile[0] = 0xDEAD;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int R = 1000, C = 128;
  std::vector<uint16_t> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 256 + c);   // r < 256 keeps it unique in 16 bits
  uint16_t *d, *out;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&out, 512);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)p;
  struct Case { const char* name; int base_off; uint64_t cols; int col; CUtensorMapSwizzle sw; } cases[] = {
      {"A full map col=0  no swizzle", 0, 128, 0, CU_TENSOR_MAP_SWIZZLE_NONE}, {"A full map col=64 no swizzle", 0, 128, 64, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"B window base+0   no swizzle", 0, 64, 0, CU_TENSOR_MAP_SWIZZLE_NONE}, {"B window base+64  no swizzle", 64, 64, 0, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"A full map col=64 SW128", 0, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B}, {"B window base+64  SW128", 64, 64, 0, CU_TENSOR_MAP_SWIZZLE_128B}};
  for (auto& cs : cases) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cs.cols, (cuuint64_t)R}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
    CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d + cs.base_off, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, cs.sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    probe<<<1, 128>>>(m, cs.col, 3, 200, 77, 5000 /* out of range */, out);
    cudaError_t e = cudaDeviceSynchronize();
    uint16_t o[256];
    cudaMemcpy(o, out, 512, cudaMemcpyDeviceToHost);
    printf("%s: encode rc=%d sync=%s\n", cs.name, (int)rc, cudaGetErrorString(e));
    for (int r = 0; r < 4; ++r) {
      printf("   smem row %d:", r);
      for (int c = 0; c < 64; c += 8) printf(" (r%d,c%d)", o[r * 64 + c] >> 8, o[r * 64 + c] & 255);
      printf("\n");
    }
  }
  return 0;
}
