// K0 and weight-layout kernels (HBM-bound data movement).
#include "common.h"
#include "optype.cuh"
#include "stream_f16.h"
#include "ptx.cuh"

namespace vnqa {

// K0: (B, C, 130, T) fp32 with T innermost -> [img][130][C] in the operand type (replaces v_input[:ct, :, :, :, i] of
// models/film_attn_pt_stem.py:210).  One block per (image row h, 128-byte channel slab, clip): for one channel the
// 13 pixels x T frames of an image row are ONE contiguous run of 13*T floats (1820 B at T = 35) in v_input, so the
// block reads 32 / 64 long runs (round 1 read 140-byte runs and reached 0.35 of the HBM roofline), converts into a
// [slab][13*T] shared tile with an odd pitch (conflict-free column reads) and writes 13*T runs of 128 contiguous bytes.
template <typename T, int CH>
__global__ void __launch_bounds__(256) frames_to_nhwc_rows_kernel(const float* __restrict__ v, T* __restrict__ x,
                                                                  const int* __restrict__ img_of, int C, int Tn,
                                                                  int pitch) {
  extern __shared__ uint8_t k0_raw[];
  T* tile = reinterpret_cast<T*>(k0_raw);
  const int h = blockIdx.x;   // image row 0..9
  const int cs = blockIdx.y;  // channel slab
  const int b = blockIdx.z;
  const int run = VNQA_IMG_W * Tn;
  const int nc = min(CH, C - cs * CH);   // the last slab may be partial (C is a multiple of 32)
  const float* src = v + (((long long)b * C + cs * CH) * VNQA_IMG_PIX + h * VNQA_IMG_W) * Tn;
  // 8 independent loads in flight per thread (the kernel is pure latency otherwise: measured 1.0 TB/s with one);
  // element i of the slab is (channel i / run, position i % run) - tracked incrementally, no division in the loop
  {
    const int total = nc * run;
    constexpr int U = 8;
    int c = threadIdx.x / run, j = threadIdx.x - c * run;
    const int dc = blockDim.x / run, dj = blockDim.x - dc * run;      // one stride of blockDim.x elements
    for (int i = threadIdx.x; i < total; i += U * blockDim.x) {
      float val[U];
      int cc[U], jj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        cc[u] = c; jj[u] = j;
        val[u] = (i + u * (int)blockDim.x < total) ? __ldg(src + (long long)c * VNQA_IMG_PIX * Tn + j) : 0.f;
        c += dc; j += dj;
        if (j >= run) { j -= run; ++c; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i + u * (int)blockDim.x < total) Op<T>::st(tile + cc[u] * pitch + jj[u], val[u], 0);
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr int PER_LANE = CH / 32;      // 1 float or 2 halves per lane: 128 B per warp store
  if (PER_LANE * lane >= nc) return;
  for (int idx = threadIdx.x >> 5; idx < run; idx += blockDim.x >> 5) {
    const int px = idx / Tn, t = idx - px * Tn;
    const int img = __ldg(img_of + b * Tn + t);
    if (img < 0) continue;
    T* dst = x + ((long long)img * VNQA_IMG_PIX + h * VNQA_IMG_W + px) * C + cs * CH + PER_LANE * lane;
    if (PER_LANE == 1) {
      dst[0] = tile[lane * pitch + idx];
    } else {
      const T a = tile[(2 * lane) * pitch + idx], bb = tile[(2 * lane + 1) * pitch + idx];
      T pair[2] = {a, bb};
      *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(pair);
    }
  }
}

template <typename T, int CH>
static int launch_k0(const float* v_input, void* x_out, const int32_t* img_of, int B, int C, int Tn, cudaStream_t st) {
  const int pitch = (VNQA_IMG_W * Tn) | 1;
  const int smem = CH * pitch * (int)sizeof(T);
  static int max_smem = 0;
  if (smem > max_smem) {
    VNQA_CHECK_CUDA(cudaFuncSetAttribute(frames_to_nhwc_rows_kernel<T, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem));
    max_smem = smem;
  }
  dim3 grid(VNQA_IMG_H, (C + CH - 1) / CH, B);
  frames_to_nhwc_rows_kernel<T, CH><<<grid, 256, smem, st>>>(v_input, reinterpret_cast<T*>(x_out), img_of, C, Tn, pitch);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// K0 for FRAME-BLOCKED input: every (clip, frame) is one contiguous (C, 130) NCHW block at v + b*stride_b + t*stride_t.
// That is what the reference's drivers hand to the model: torch.stack(per-frame stem outputs).permute(1,2,3,4,0)
// followed by the length sort v_inputs[perm_idx] (eval/q_and_v_eval.py:110-116) is a (B, C, 10, 13, T)-shaped VIEW of
// memory ordered (B, T, C, 10, 13).  Reading it directly saves the .contiguous() pass (596 MB read + written) that the
// T-innermost kernel above would need.  One block per (channel slab, frame image): CH*130 contiguous floats in, 130
// rows of 128 bytes out.
template <typename T, int CH>
__global__ void __launch_bounds__(256) frames_blocked_to_nhwc_kernel(const float* __restrict__ v, T* __restrict__ x,
                                                                     const int* __restrict__ img_of, int C, int Tn,
                                                                     long long stride_b, long long stride_t) {
  constexpr int PITCH = VNQA_IMG_PIX + 1;
  __shared__ float tile[CH * PITCH];
  const int cs = blockIdx.x, bt = blockIdx.y;
  const int img = __ldg(img_of + bt);
  if (img < 0) return;
  const int b = bt / Tn, t = bt - b * Tn;
  const int nc = min(CH, C - cs * CH);
  const float* src = v + b * stride_b + t * stride_t + (long long)cs * CH * VNQA_IMG_PIX;
  const int total = nc * VNQA_IMG_PIX;
  constexpr int U = 4;
  for (int i0 = threadIdx.x; i0 < total; i0 += U * 256) {
    float val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) val[u] = (i0 + u * 256 < total) ? __ldg(src + i0 + u * 256) : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 256;
      if (i < total) {
        const int c = i / VNQA_IMG_PIX;
        tile[c * PITCH + (i - c * VNQA_IMG_PIX)] = val[u];
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr int PER_LANE = CH / 32;
  if (PER_LANE * lane >= nc) return;
  T* dst0 = x + (long long)img * VNQA_IMG_PIX * C + cs * CH + PER_LANE * lane;
  for (int p = threadIdx.x >> 5; p < VNQA_IMG_PIX; p += 8) {
    T* dst = dst0 + (long long)p * C;
    if (PER_LANE == 1) {
      Op<T>::st(dst, tile[lane * PITCH + p], 0);
    } else {
      T pair[2];
      Op<T>::st(&pair[0], tile[(2 * lane) * PITCH + p], 0);
      Op<T>::st(&pair[1], tile[(2 * lane + 1) * PITCH + p], 0);
      *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(pair);
    }
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

extern "C" int vnqa_frames_to_nhwc(const float* v_input, void* x_out, const int32_t* img_of, int32_t B, int32_t C,
                                   int32_t T, int32_t act_dtype, void* stream) {
  VNQA_REQUIRE(v_input && x_out && img_of, "vnqa_frames_to_nhwc: null pointer");
  VNQA_REQUIRE(C % 32 == 0 && C > 0, "vnqa_frames_to_nhwc: C must be a multiple of 32 (got %d)", C);
  VNQA_REQUIRE(T > 0 && T <= 64 && B > 0 && B <= 65535, "vnqa_frames_to_nhwc: bad B/T (%d, %d)", B, T);
  return act_dtype == VNQA_ACT_FP16 ? launch_k0<__half, 64>(v_input, x_out, img_of, B, C, T, as_stream(stream))
                                    : launch_k0<float, 32>(v_input, x_out, img_of, B, C, T, as_stream(stream));
}

extern "C" int vnqa_frames_blocked_to_nhwc(const float* v_input, void* x_out, const int32_t* img_of, int32_t B, int32_t C,
                                           int32_t T, int64_t stride_b, int64_t stride_t, int32_t act_dtype,
                                           void* stream) {
  VNQA_REQUIRE(v_input && x_out && img_of, "vnqa_frames_blocked_to_nhwc: null pointer");
  VNQA_REQUIRE(C % 32 == 0 && C > 0, "vnqa_frames_blocked_to_nhwc: C must be a multiple of 32 (got %d)", C);
  VNQA_REQUIRE(T > 0 && B > 0 && (long long)B * T <= 65535, "vnqa_frames_blocked_to_nhwc: bad B/T (%d, %d)", B, T);
  if (act_dtype == VNQA_ACT_FP16) {
    dim3 grid((C + 63) / 64, B * T);
    frames_blocked_to_nhwc_kernel<__half, 64><<<grid, 256, 0, as_stream(stream)>>>(
        v_input, static_cast<__half*>(x_out), img_of, C, T, stride_b, stride_t);
  } else {
    dim3 grid((C + 31) / 32, B * T);
    frames_blocked_to_nhwc_kernel<float, 32><<<grid, 256, 0, as_stream(stream)>>>(
        v_input, static_cast<float*>(x_out), img_of, C, T, stride_b, stride_t);
  }
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
