// tcgen05 implicit-GEMM sparse convolution (placeholder until the kernel lands: reports "unsupported" so the
// dispatcher keeps using the SIMT kernels).
#pragma once
#include "common.cuh"

namespace b2pc {
inline bool spconv_umma_supported(int, int, int) { return false; }
inline int launch_gather_gemm_umma(const void*, const void*, const void*, const int32_t*, int64_t, int64_t, int64_t, int, int, int,
                                   int, int, int, void*, cudaStream_t) {
  set_error("spconv_gather_gemm: tcgen05 kernel not built");
  return B2PC_ERR_UNSUPPORTED;
}
}  // namespace b2pc
