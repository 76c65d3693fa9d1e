// HBM-bound kernels that touch trunk activations, fp16 storage (VNQA_ACT_FP16): bn_apply,
// bn_backward, film_backward, weight packing and fp32 -> fp16 conversion.  Same arithmetic as their fp32 namesakes in
// layout.cu / small.cu (fp32 math, fp16 loads/stores, 16-byte accesses); reached through the same C entry points,
// which dispatch on act_dtype.  Backward signals are stored multiplied by a power-of-two loss scale; everything that
// leaves the fp16 domain (parameter gradients) is multiplied by grad_unscale = 1 / scale.
#include <cuda_fp16.h>

#include "common.h"
#include "stream_f16.h"

namespace vnqa {

struct alignas(16) Half8 {
  __half2 a, b, c, d;
};
__device__ __forceinline__ void unpack8(const Half8& h, float* f) {
  const float2 p0 = __half22float2(h.a), p1 = __half22float2(h.b), p2 = __half22float2(h.c), p3 = __half22float2(h.d);
  f[0] = p0.x; f[1] = p0.y; f[2] = p1.x; f[3] = p1.y; f[4] = p2.x; f[5] = p2.y; f[6] = p3.x; f[7] = p3.y;
}
__device__ __forceinline__ Half8 pack8(const float* f) {
  Half8 h;
  h.a = __floats2half2_rn(f[0], f[1]); h.b = __floats2half2_rn(f[2], f[3]);
  h.c = __floats2half2_rn(f[4], f[5]); h.d = __floats2half2_rn(f[6], f[7]);
  return h;
}
// one 16-byte store (a plain struct assignment is split into four 4-byte stores: 25 % sector efficiency)
__device__ __forceinline__ void st8(Half8* p, const Half8& h) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&h);
}
__device__ __forceinline__ Half8 ldg8(const Half8* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  return *reinterpret_cast<const Half8*>(&u);
}

// x[row][c] = fp16(y[row][c] * scale[frame(img)][c] + shift[frame(img)][c]); 8 channels per thread
__global__ void __launch_bounds__(256) bn_apply_h_kernel(const float4* __restrict__ y, Half8* __restrict__ x,
                                                         const int* __restrict__ img_frame,
                                                         const float4* __restrict__ scale,
                                                         const float4* __restrict__ shift, long long n_vec, int C8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const long long row = i / C8;
    const int f = __ldg(img_frame + row / VNQA_IMG_PIX);
    const float4 v0 = y[2 * i], v1 = y[2 * i + 1];
    const float4 s0 = __ldg(scale + ((long long)f * C8 + c8) * 2), s1 = __ldg(scale + ((long long)f * C8 + c8) * 2 + 1);
    const float4 t0 = __ldg(shift + ((long long)f * C8 + c8) * 2), t1 = __ldg(shift + ((long long)f * C8 + c8) * 2 + 1);
    float o[8] = {fmaf(v0.x, s0.x, t0.x), fmaf(v0.y, s0.y, t0.y), fmaf(v0.z, s0.z, t0.z), fmaf(v0.w, s0.w, t0.w),
                  fmaf(v1.x, s1.x, t1.x), fmaf(v1.y, s1.y, t1.y), fmaf(v1.z, s1.z, t1.z), fmaf(v1.w, s1.w, t1.w)};
    st8(x + i, pack8(o));
  }
}

// BatchNorm backward, pass 1: per (frame, channel) s1 = sum dx, s2 = sum dx * xhat.  dx is fp16 (loss-scaled), y fp32.
// One block per (image, 256-channel slab): thread = channel pair.
__global__ void __launch_bounds__(128) bn_bwd_reduce_h_kernel(const __half* __restrict__ dx, const float* __restrict__ y,
                                                              const int* __restrict__ img_frame,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ rstd, int C,
                                                              float* __restrict__ s1, float* __restrict__ s2) {
  const int img = blockIdx.y;
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const int f = img_frame[img];
  const float2 m = *reinterpret_cast<const float2*>(mean + (long long)f * C + c);
  const float2 r = *reinterpret_cast<const float2*>(rstd + (long long)f * C + c);
  const __half* dp = dx + (long long)img * VNQA_IMG_PIX * C + c;
  const float* yp = y + (long long)img * VNQA_IMG_PIX * C + c;
  float a1x = 0.f, a1y = 0.f, a2x = 0.f, a2y = 0.f;
#pragma unroll 10
  for (int p = 0; p < VNQA_IMG_PIX; ++p) {
    const float2 d = __half22float2(__ldg(reinterpret_cast<const __half2*>(dp + (long long)p * C)));
    const float2 yy = __ldg(reinterpret_cast<const float2*>(yp + (long long)p * C));
    a1x += d.x; a1y += d.y;
    a2x = fmaf(d.x, (yy.x - m.x) * r.x, a2x);
    a2y = fmaf(d.y, (yy.y - m.y) * r.y, a2y);
  }
  atomicAdd(s1 + (long long)f * C + c, a1x); atomicAdd(s1 + (long long)f * C + c + 1, a1y);
  atomicAdd(s2 + (long long)f * C + c, a2x); atomicAdd(s2 + (long long)f * C + c + 1, a2y);
}

// pass 2: dy = fp16(gamma*rstd*(dx - s1/n - xhat*s2/n)) masked by the ReLU before BN (y > 0); per-channel sum of dy
// (conv_init bias gradient, unscaled).
__global__ void __launch_bounds__(128) bn_bwd_apply_h_kernel(const __half* __restrict__ dx, const float* __restrict__ y,
                                                             const int* __restrict__ img_frame,
                                                             const int* __restrict__ frame_ct,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ s1, const float* __restrict__ s2,
                                                             int C, __half* __restrict__ dy, float* __restrict__ dbias,
                                                             float grad_unscale) {
  const int img = blockIdx.y;
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const int f = img_frame[img];
  const float inv_n = 1.f / ((float)frame_ct[f] * VNQA_IMG_PIX);
  const long long fc = (long long)f * C + c;
  const float2 m = *reinterpret_cast<const float2*>(mean + fc), r = *reinterpret_cast<const float2*>(rstd + fc);
  const float2 gm = *reinterpret_cast<const float2*>(gamma + c);
  const float gx = gm.x * r.x, gy = gm.y * r.y;
  const float k1x = s1[fc] * inv_n, k1y = s1[fc + 1] * inv_n, k2x = s2[fc] * inv_n, k2y = s2[fc + 1] * inv_n;
  const long long base = (long long)img * VNQA_IMG_PIX * C + c;
  float accx = 0.f, accy = 0.f;
#pragma unroll 10
  for (int p = 0; p < VNQA_IMG_PIX; ++p) {
    const long long i = base + (long long)p * C;
    const float2 yy = __ldg(reinterpret_cast<const float2*>(y + i));
    const float2 d = __half22float2(__ldg(reinterpret_cast<const __half2*>(dx + i)));
    const float ox = yy.x > 0.f ? gx * (d.x - k1x - (yy.x - m.x) * r.x * k2x) : 0.f;
    const float oy = yy.y > 0.f ? gy * (d.y - k1y - (yy.y - m.y) * r.y * k2y) : 0.f;
    const __half2 o = __floats2half2_rn(ox, oy);
    *reinterpret_cast<__half2*>(dy + i) = o;
    const float2 of = __half22float2(o);
    accx += of.x; accy += of.y;
  }
  atomicAdd(dbias + c, accx * grad_unscale);
  atomicAdd(dbias + c + 1, accy * grad_unscale);
}

__global__ void bn_bwd_params_h_kernel(const float* __restrict__ s1, const float* __restrict__ s2, int n_frames, int C,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float grad_unscale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int f = 0; f < n_frames; ++f) {
    a += s2[(long long)f * C + c];
    b += s1[(long long)f * C + c];
  }
  dgamma[c] += a * grad_unscale;
  dbeta[c] += b * grad_unscale;
}

// FiLM backward (autograd of models/film_attn_pt_stem.py:229-239) on fp16 dOut / z, producing fp16 dz:
//   dF = dOut * [alpha*z+beta > 0]; dz = alpha*dF; dalpha[img][c] = unscale * sum_pix dF*z; dbeta likewise;
//   dbias[c] += unscale * sum dz.  One block per image; a thread owns 8 consecutive channels (16-byte accesses).
__global__ void __launch_bounds__(128) film_bwd_h_kernel(const Half8* __restrict__ dout, const Half8* __restrict__ z,
                                                         const float* __restrict__ film,
                                                         const int* __restrict__ film_row, int film_ld, int alpha_off,
                                                         int C, Half8* __restrict__ dz, float* __restrict__ dfilm,
                                                         float* __restrict__ dbias, int mask_relu, float grad_unscale) {
  const int img = blockIdx.x;
  const int C8 = C >> 3;
  const long long frow = (long long)film_row[img] * film_ld + alpha_off;
  for (int q = threadIdx.x; q < C8; q += blockDim.x) {
    float al[8], be[8], da[8], db[8], ds[8];
    {
      const float4 a0 = *reinterpret_cast<const float4*>(film + frow + 8 * q);
      const float4 a1 = *reinterpret_cast<const float4*>(film + frow + 8 * q + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(film + frow + C + 8 * q);
      const float4 b1 = *reinterpret_cast<const float4*>(film + frow + C + 8 * q + 4);
      al[0] = a0.x; al[1] = a0.y; al[2] = a0.z; al[3] = a0.w; al[4] = a1.x; al[5] = a1.y; al[6] = a1.z; al[7] = a1.w;
      be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { da[j] = 0.f; db[j] = 0.f; ds[j] = 0.f; }
    const Half8* zp = z + (long long)img * VNQA_IMG_PIX * C8 + q;
    const Half8* dp = dout + (long long)img * VNQA_IMG_PIX * C8 + q;
    Half8* op = dz + (long long)img * VNQA_IMG_PIX * C8 + q;
#pragma unroll 2
    for (int p = 0; p < VNQA_IMG_PIX; ++p) {
      float zz[8], dd[8], o[8];
      unpack8(ldg8(zp + (long long)p * C8), zz);
      unpack8(ldg8(dp + (long long)p * C8), dd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = fmaf(al[j], zz[j], be[j]) > 0.f ? dd[j] : 0.f;
        da[j] = fmaf(d, zz[j], da[j]);
        db[j] += d;
        o[j] = al[j] * d;
      }
      const Half8 h = pack8(o);
      st8(op + (long long)p * C8, h);
      float back[8];
      unpack8(h, back);
#pragma unroll
      for (int j = 0; j < 8; ++j) ds[j] += back[j];
    }
    // each (film row, channel) is owned by exactly one image: plain stores.  When the FiLM generator ends in a ReLU
    // (:179, mask_relu) its gradient is masked where the output is 0
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dfilm[frow + 8 * q + j] = (!mask_relu || al[j] > 0.f) ? da[j] * grad_unscale : 0.f;
      dfilm[frow + C + 8 * q + j] = (!mask_relu || be[j] > 0.f) ? db[j] * grad_unscale : 0.f;
      atomicAdd(dbias + 8 * q + j, ds[j] * grad_unscale);
    }
  }
}


// (Cout, Cin, taps) fp32 -> fp16 [tap][rows_pad][Cin], or the mirrored / transposed data-gradient operand
__global__ void pack_conv_weight_h_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin,
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
    out[i] = __float2half_rn(val);
  }
}

__global__ void to_half_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(__ldg(in + i) * scale);
}

// ------------------------------------------------------------------------------------------------ host side
int pack_conv_weight_h(const float* w, void* packed, int cout, int cin, int taps, int rows_pad, int transpose_flip,
                       cudaStream_t st) {
  const long long total = (long long)taps * rows_pad * (transpose_flip ? cout : cin);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  pack_conv_weight_h_kernel<<<blocks, 256, 0, st>>>(w, reinterpret_cast<__half*>(packed), cout, cin, taps, rows_pad,
                                                    transpose_flip);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int bn_apply_h(const float* y, void* x, const int32_t* img_frame, const float* scale, const float* shift, int n_img,
               int C, cudaStream_t st) {
  VNQA_REQUIRE(C % 8 == 0, "vnqa_bn_apply: C must be a multiple of 8 for fp16 activations");
  const long long n_vec = (long long)n_img * VNQA_IMG_PIX * C / 8;
  const int blocks = (int)((n_vec + 255) / 256 < 148 * 8 ? (n_vec + 255) / 256 : 148 * 8);
  bn_apply_h_kernel<<<blocks, 256, 0, st>>>((const float4*)y, (Half8*)x, img_frame, (const float4*)scale,
                                            (const float4*)shift, n_vec, C / 8);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int bn_backward_h(const void* dx, const float* y, const int32_t* img_frame, const int32_t* frame_ct, int n_frames,
                  const float* mean, const float* rstd, const float* gamma, int n_img, int C, float* s1, float* s2,
                  void* dy, float* dbias, float* dgamma, float* dbeta, float grad_unscale, cudaStream_t st) {
  VNQA_REQUIRE(C % 2 == 0, "vnqa_bn_backward: C must be even for fp16 activations");
  VNQA_CHECK_CUDA(cudaMemsetAsync(s1, 0, (size_t)n_frames * C * 4, st));
  VNQA_CHECK_CUDA(cudaMemsetAsync(s2, 0, (size_t)n_frames * C * 4, st));
  dim3 grid((C / 2 + 127) / 128, n_img);
  bn_bwd_reduce_h_kernel<<<grid, 128, 0, st>>>((const __half*)dx, y, img_frame, mean, rstd, C, s1, s2);
  bn_bwd_apply_h_kernel<<<grid, 128, 0, st>>>((const __half*)dx, y, img_frame, frame_ct, mean, rstd, gamma, s1, s2, C,
                                              (__half*)dy, dbias, grad_unscale);
  bn_bwd_params_h_kernel<<<(C + 127) / 128, 128, 0, st>>>(s1, s2, n_frames, C, dgamma, dbeta, grad_unscale);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int film_backward_h(const void* dout, const void* z, const float* film, const int32_t* film_row, int film_ld,
                    int alpha_off, int n_img, int C, void* dz, float* dfilm, float* dbias, int mask_relu,
                    float grad_unscale, cudaStream_t st) {
  VNQA_REQUIRE(C % 8 == 0 && film_ld % 4 == 0 && alpha_off % 4 == 0,
               "vnqa_film_backward: fp16 activations need C multiple of 8, film_ld/alpha_off multiples of 4");
  const int threads = C / 8 >= 128 ? 128 : ((C / 8 + 31) / 32) * 32;
  film_bwd_h_kernel<<<n_img, threads, 0, st>>>((const Half8*)dout, (const Half8*)z, film, film_row, film_ld, alpha_off,
                                               C, (Half8*)dz, dfilm, dbias, mask_relu, grad_unscale);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vnqa

extern "C" int vnqa_to_half(const float* in, void* out, int64_t n, float scale, void* stream) {
  VNQA_REQUIRE(in && out && n >= 0, "vnqa_to_half: bad arguments");
  if (n == 0) return 0;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  vnqa::to_half_kernel<<<blocks, 256, 0, vnqa::as_stream(stream)>>>(in, reinterpret_cast<__half*>(out), n, scale);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
