// K0 and weight-layout kernels (HBM-bound data movement).
#include "common.h"
#include "stream_f16.h"
#include "ptx.cuh"

namespace vnqa {

// (B, C, 130, T) with T innermost -> [img][130][C].  One block per (clip, pixel, 32-channel slab): the slab
// [32 c][T] is read as 32 runs of T contiguous floats, transposed through shared memory and written as T runs of
// 32 contiguous channels (128 B).  Replaces v_input[:ct, :, :, :, i] of models/film_attn_pt_stem.py:210.
__global__ void __launch_bounds__(256) frames_to_nhwc_kernel(const float* __restrict__ v, float* __restrict__ x,
                                                             const int* __restrict__ img_of, int C, int T) {
  __shared__ float tile[32][65];
  const int p = blockIdx.x;   // pixel 0..129
  const int cs = blockIdx.y;  // channel slab
  const int b = blockIdx.z;
  const float* src = v + (((long long)b * C + cs * 32) * VNQA_IMG_PIX + p) * T;
  for (int i = threadIdx.x; i < 32 * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    tile[c][t] = __ldg(src + (long long)c * VNQA_IMG_PIX * T + t);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x >> 5; t < T; t += blockDim.x >> 5) {
    const int img = __ldg(img_of + b * T + t);
    if (img >= 0) x[((long long)img * VNQA_IMG_PIX + p) * C + cs * 32 + lane] = tf32_round(tile[lane][t]);
  }
}

// (Cout, Cin, taps) -> [tap][rows_pad][Cin]   (forward operand), or
// (Cout, Cin, taps) -> [taps-1-tap][rows_pad >= Cin][Cout]   (data-gradient operand: mirrored taps, in/out swapped)
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin,
                                        int taps, int rows_pad, int transpose_flip) {
  const long long K = transpose_flip ? cout : cin;
  const long long total = (long long)taps * rows_pad * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long r2 = i / K;
    const int row = (int)(r2 % rows_pad);
    const int tap = (int)(r2 / rows_pad);
    float val = 0.f;
    if (!transpose_flip) {
      if (row < cout) val = __ldg(w + ((long long)row * cin + k) * taps + tap);
    } else {
      if (row < cin) val = __ldg(w + ((long long)k * cin + row) * taps + (taps - 1 - tap));
    }
    out[i] = tf32_round(val);
  }
}

// Linear weight (n, C*130) col = c*130 + p  <->  (n_pad, 130*C) col = p*C + c
__global__ void permute_fc_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int n, int n_pad, int C,
                                         int inverse) {
  const long long K = (long long)C * VNQA_IMG_PIX;
  const long long total = (long long)(inverse ? n : n_pad) * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long col = i % K;
    const int row = (int)(i / K);
    if (!inverse) {
      const int pix = (int)(col / C), c = (int)(col % C);  // output col = p*C + c
      out[i] = row < n ? tf32_round(__ldg(w + (long long)row * K + (long long)c * VNQA_IMG_PIX + pix)) : 0.f;
    } else {
      const int c = (int)(col / VNQA_IMG_PIX), pix = (int)(col % VNQA_IMG_PIX);  // output col = c*130 + p
      out[i] = __ldg(w + (long long)row * K + (long long)pix * C + c);
    }
  }
}

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_frames_to_nhwc(const float* v_input, void* x_out_, const int32_t* img_of, int32_t B, int32_t C,
                                   int32_t T, int32_t act_dtype, void* stream) {
  VNQA_REQUIRE(v_input && x_out_ && img_of, "vnqa_frames_to_nhwc: null pointer");
  if (act_dtype == VNQA_ACT_FP16) return frames_to_nhwc_h(v_input, x_out_, img_of, B, C, T, as_stream(stream));
  float* x_out = static_cast<float*>(x_out_);
  VNQA_REQUIRE(C % 32 == 0 && C > 0, "vnqa_frames_to_nhwc: C must be a multiple of 32 (got %d)", C);
  VNQA_REQUIRE(T > 0 && T <= 64 && B > 0 && B <= 65535, "vnqa_frames_to_nhwc: bad B/T (%d, %d)", B, T);
  dim3 grid(VNQA_IMG_PIX, C / 32, B);
  frames_to_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(v_input, x_out, img_of, C, T);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_pack_conv_weight(const float* w, void* packed_, int32_t cout, int32_t cin, int32_t taps,
                                     int32_t rows_pad, int32_t transpose_flip, int32_t act_dtype, void* stream) {
  VNQA_REQUIRE(w && packed_, "vnqa_pack_conv_weight: null pointer");
  VNQA_REQUIRE(taps == 1 || taps == 9, "vnqa_pack_conv_weight: taps must be 1 or 9");
  VNQA_REQUIRE(rows_pad >= (transpose_flip ? cin : cout), "vnqa_pack_conv_weight: rows_pad too small");
  if (act_dtype == VNQA_ACT_FP16)
    return pack_conv_weight_h(w, packed_, cout, cin, taps, rows_pad, transpose_flip, as_stream(stream));
  float* packed = static_cast<float*>(packed_);
  const long long total = (long long)taps * rows_pad * (transpose_flip ? cout : cin);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  pack_conv_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, packed, cout, cin, taps, rows_pad, transpose_flip);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_permute_fc_weight(const float* w, float* out, int32_t n, int32_t n_pad, int32_t C, int32_t inverse,
                                      void* stream) {
  VNQA_REQUIRE(w && out && n_pad >= n, "vnqa_permute_fc_weight: bad arguments");
  const long long total = (long long)(inverse ? n : n_pad) * C * VNQA_IMG_PIX;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  permute_fc_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, out, n, n_pad, C, inverse);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
