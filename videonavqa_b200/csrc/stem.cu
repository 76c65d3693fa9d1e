// Layout kernels at the two ends of the frozen stem (models/obj_detector.py:69-86, called per frame by
// eval/q_and_v_eval.py:104-110).  The convolutions in between are conv_gemm.cu with the 40 x 52 / 20 x 26 / 10 x 13
// geometries; these two kernels are pure HBM data movement: 64-channel x 32-pixel tiles through shared memory so that
// both the NCHW side (runs of 32 pixels = 128 B) and the NHWC side (runs of 64 channels = 128 / 256 B) are coalesced.
#include "common.h"
#include "optype.cuh"
#include "ptx.cuh"

namespace vnqa {

constexpr int ST_C = 64;   // channels per tile
constexpr int ST_P = 32;   // pixels per tile

// x (N, C, P) fp32  ->  out [N][P][C] in the operand type, out = x * scale[c] + shift[c] (bn_input in eval mode,
// obj_detector.py:70; the affine cannot be folded into conv11 because the zero padding is applied AFTER it)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_affine_kernel(const float* __restrict__ x, T* __restrict__ out,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, int C, int P) {
  __shared__ float tile[ST_C][ST_P + 1];
  const int p0 = blockIdx.x * ST_P, c0 = blockIdx.y * ST_C;
  const long long n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
#pragma unroll
  for (int i = 0; i < ST_C / 8; ++i) {
    const int c = c0 + ty + 8 * i, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < P) {
      v = __ldg(x + (n * C + c) * P + p);
      if (scale != nullptr) v = fmaf(v, __ldg(scale + c), __ldg(shift + c));
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
  // each warp writes one pixel row at a time: lane -> channels 2*lane, 2*lane+1
#pragma unroll
  for (int i = 0; i < ST_P / 8; ++i) {
    const int pl = ty + 8 * i, p = p0 + pl;
    if (p >= P) continue;
    T* dst = out + (n * P + p) * C + c0;
    const int c = 2 * tx;
    if (c0 + c + 1 < C) {
      if (Op<T>::HALF) {
        T pair[2];
        Op<T>::st(&pair[0], tile[c][pl], 0);
        Op<T>::st(&pair[1], tile[c + 1][pl], 0);
        *reinterpret_cast<uint32_t*>(dst + c) = *reinterpret_cast<const uint32_t*>(pair);
      } else {
        Op<T>::st(dst + c, tile[c][pl], 0);
        Op<T>::st(dst + c + 1, tile[c + 1][pl], 0);
      }
    } else if (c0 + c < C) {
      Op<T>::st(dst + c, tile[c][pl], 0);
    }
  }
}

// Wide variant for P % 4 == 0 (the stem's 40 x 52 = 2080-pixel input): 64 channels x 64 pixels per block, 16-byte loads
// (4 independent ones per thread = 16 KB in flight per block, 256-byte contiguous runs per channel), one warp per
// output pixel row (64 channels = 128 B of half / 256 B of fp32).  The narrow kernel above reached 1.7-2.1 TB/s.
constexpr int SW_P = 64;
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_affine_wide_kernel(const float* __restrict__ x, T* __restrict__ out,
                                                                       const float* __restrict__ scale,
                                                                       const float* __restrict__ shift, int C, int P) {
  __shared__ float tile[ST_C][SW_P + 1];
  const int p0 = blockIdx.x * SW_P, c0 = blockIdx.y * ST_C;
  const long long n = blockIdx.z;
  const int q = threadIdx.x & 15, r = threadIdx.x >> 4;        // 16 float4 per channel row, 16 channel rows per pass
  float4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + r + 16 * i, p = p0 + 4 * q;
    v[i] = (c < C && p < P) ? __ldg(reinterpret_cast<const float4*>(x + (n * C + c) * P + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = r + 16 * i, c = c0 + cl;
    float sc = 1.f, sh = 0.f;
    if (scale != nullptr && c < C) { sc = __ldg(scale + c); sh = __ldg(shift + c); }
    tile[cl][4 * q + 0] = fmaf(v[i].x, sc, sh);
    tile[cl][4 * q + 1] = fmaf(v[i].y, sc, sh);
    tile[cl][4 * q + 2] = fmaf(v[i].z, sc, sh);
    tile[cl][4 * q + 3] = fmaf(v[i].w, sc, sh);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = 2 * lane;
#pragma unroll
  for (int i = 0; i < SW_P / 8; ++i) {
    const int pl = w + 8 * i, p = p0 + pl;
    if (p >= P) continue;
    T* dst = out + (n * P + p) * C + c0;
    if (c0 + c + 1 < C) {
      if (Op<T>::HALF) {
        T pair[2];
        Op<T>::st(&pair[0], tile[c][pl], 0);
        Op<T>::st(&pair[1], tile[c + 1][pl], 0);
        *reinterpret_cast<uint32_t*>(dst + c) = *reinterpret_cast<const uint32_t*>(pair);
      } else {
        Op<T>::st(dst + c, tile[c][pl], 0);
        Op<T>::st(dst + c + 1, tile[c + 1][pl], 0);
      }
    } else if (c0 + c < C) {
      Op<T>::st(dst + c, tile[c][pl], 0);
    }
  }
}

// x [N][P][C] (fp32 or half)  ->  out (N, C, P) fp32: the layout the reference's callers expect back
// (torch.stack(v_features).permute(1, 2, 3, 4, 0), q_and_v_eval.py:110)
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ out, int C,
                                                           int P) {
  __shared__ float tile[ST_P][ST_C + 1];
  const int p0 = blockIdx.x * ST_P, c0 = blockIdx.y * ST_C;
  const long long n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < ST_P / 8; ++i) {
    const int pl = ty + 8 * i, p = p0 + pl;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = tx + 32 * h;
      tile[pl][c] = (p < P && c0 + c < C) ? Op<T>::ld(x + (n * P + p) * C + c0 + c) : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ST_C / 8; ++i) {
    const int cl = ty + 8 * i, c = c0 + cl, p = p0 + tx;
    if (c < C && p < P) out[(n * C + c) * P + p] = tile[tx][cl];
  }
}

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_nchw_to_nhwc_affine(const float* x, void* out, const float* scale, const float* shift,
                                        int32_t n_img, int32_t C, int32_t P, int32_t act_dtype, void* stream) {
  VNQA_REQUIRE(x && out, "vnqa_nchw_to_nhwc_affine: null pointer");
  VNQA_REQUIRE((scale == nullptr) == (shift == nullptr), "vnqa_nchw_to_nhwc_affine: scale and shift go together");
  VNQA_REQUIRE(n_img > 0 && n_img <= 65535 && C > 0 && C % 2 == 0 && P > 0, "vnqa_nchw_to_nhwc_affine: bad sizes (%d, %d, %d)",
               n_img, C, P);
  if (P % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 grid((P + SW_P - 1) / SW_P, (C + ST_C - 1) / ST_C, n_img);
    if (act_dtype == VNQA_ACT_FP16)
      nchw_to_nhwc_affine_wide_kernel<__half><<<grid, 256, 0, as_stream(stream)>>>(x, static_cast<__half*>(out), scale, shift, C, P);
    else
      nchw_to_nhwc_affine_wide_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(x, static_cast<float*>(out), scale, shift, C, P);
    VNQA_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid((P + ST_P - 1) / ST_P, (C + ST_C - 1) / ST_C, n_img);
  if (act_dtype == VNQA_ACT_FP16)
    nchw_to_nhwc_affine_kernel<__half><<<grid, 256, 0, as_stream(stream)>>>(x, static_cast<__half*>(out), scale, shift, C, P);
  else
    nchw_to_nhwc_affine_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(x, static_cast<float*>(out), scale, shift, C, P);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_nhwc_to_nchw(const void* x, float* out, int32_t n_img, int32_t C, int32_t P, int32_t x_half,
                                 void* stream) {
  VNQA_REQUIRE(x && out, "vnqa_nhwc_to_nchw: null pointer");
  VNQA_REQUIRE(n_img > 0 && n_img <= 65535 && C > 0 && P > 0, "vnqa_nhwc_to_nchw: bad sizes (%d, %d, %d)", n_img, C, P);
  dim3 grid((P + ST_P - 1) / ST_P, (C + ST_C - 1) / ST_C, n_img);
  if (x_half)
    nhwc_to_nchw_kernel<__half><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __half*>(x), out, C, P);
  else
    nhwc_to_nchw_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), out, C, P);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
