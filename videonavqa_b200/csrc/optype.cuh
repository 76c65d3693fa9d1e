// Operand-type traits of the tensor-core kernels: ONE kernel source serves both storage types.
//   Op<float>  : fp32 storage holding tf32-representable values  -> tcgen05.mma.kind::tf32 (32 elements / 128-B row)
//   Op<__half> : fp16 storage (same 10-bit mantissa as tf32)     -> tcgen05.mma.kind::f16  (64 elements / 128-B row)
// Accumulation is fp32 in TMEM in both cases.  The tensor core TRUNCATES fp32 operands of kind::tf32, so every fp32
// tensor that feeds an MMA is rounded to nearest once by its producer (st / round below); fp16 stores round to nearest
// even by construction.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace vnqa {

template <typename T>
struct Op;

template <>
struct Op<float> {
  static constexpr bool HALF = false;
  static constexpr int BYTES = 4;
  static constexpr int BK = 32;        // reduction elements per 128-byte stage row
  static constexpr int UMMA_K = 8;     // reduction elements per MMA (32 bytes)
  __host__ __device__ static constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return umma_idesc_tf32(M, N, a_mn, b_mn);
  }
  template <int NCTA>
  __device__ __forceinline__ static void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (NCTA == 1) umma_tf32(d, a, b, idesc, acc);
    else umma_tf32_2sm(d, a, b, idesc, acc);
  }
  __device__ __forceinline__ static float ld(const float* p) { return __ldg(p); }
  __device__ __forceinline__ static float round(float x) { return tf32_round(x); }
  // exact != 0: keep the full fp32 value (the tensor does not feed an MMA)
  __device__ __forceinline__ static void st(float* p, float x, int exact) { *p = exact ? x : tf32_round(x); }
};

template <>
struct Op<__half> {
  static constexpr bool HALF = true;
  static constexpr int BYTES = 2;
  static constexpr int BK = 64;
  static constexpr int UMMA_K = 16;
  __host__ __device__ static constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return umma_idesc_f16(M, N, a_mn, b_mn);
  }
  template <int NCTA>
  __device__ __forceinline__ static void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (NCTA == 1) umma_f16(d, a, b, idesc, acc);
    else umma_f16_2sm(d, a, b, idesc, acc);
  }
  __device__ __forceinline__ static float ld(const __half* p) { return __half2float(__ldg(p)); }
  __device__ __forceinline__ static float round(float x) { return __half2float(__float2half_rn(x)); }
  __device__ __forceinline__ static void st(__half* p, float x, int) { *p = __float2half_rn(x); }
};

}  // namespace vnqa
