// Shared by the fp16 candidate kernels: tcgen05 kind::f16 wrappers and the fp16 tensor-map helper.
#pragma once
#include <cuda_fp16.h>

#include "../../videonavqa_b200/csrc/common.h"
#include "../../videonavqa_b200/csrc/ptx.cuh"
#include "vnqa_f16.h"

namespace vnqa {
namespace f16 {

// D (+)= A * B, fp16 operands from shared memory, fp32 accumulate in TMEM; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor for kind::f16 with F16 A/B (format 0) and F32 D:
//   [4,6) D format 1 = F32   [7,10) A format 0 = F16   [10,13) B format 0 = F16
//   [15] A major   [16] B major (0 = K)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}

// fp16 tensor map, 128-byte swizzle (16-byte atoms: valid for K-major and MN-major 16-bit operands), zero OOB fill
int make_tmap_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

}  // namespace f16
}  // namespace vnqa
