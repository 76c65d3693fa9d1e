// Small fp32 kernels around the convolution stack: a strided SGEMM (full fp32 FFMA - the question/FiLM and
// classifier linears are tiny and stay exact), BatchNorm statistics/affine/backward, element-wise FiLM backward,
// split-K reduction, global max pooling.
#include "common.h"
#include "stream_f16.h"
#include "ptx.cuh"

namespace vnqa {

// ------------------------------------------------------------------------------------------------ SGEMM
// C[m][n] (ldc) = (accumulate ? C : 0) + sum_k A(m,k) * B(n,k) + bias[n], optional ReLU.
// A(m,k) = a[m*a_rs + k*a_cs], B(n,k) = b[n*b_rs + k*b_cs]  (any of NN/NT/TN/TT through strides).
constexpr int SG_T = 64, SG_K = 16;
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                    float* __restrict__ c, const float* __restrict__ bias, int M, int N,
                                                    int K, long long a_rs, long long a_cs, long long b_rs,
                                                    long long b_cs, long long ldc, int relu, int accumulate,
                                                    int k_per_split) {
  __shared__ float As[SG_K][SG_T + 4];
  __shared__ float Bs[SG_K][SG_T + 4];
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const bool split = gridDim.z > 1;
  K = k_end;
  for (int k0 = k_begin; k0 < k_end; k0 += SG_K) {
    for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
      int mm, kk;
      if (a_cs == 1) { kk = i % SG_K; mm = i / SG_K; } else { mm = i % SG_T; kk = i / SG_T; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? __ldg(a + gm * a_rs + gk * a_cs) : 0.f;
      int nn, k2;
      if (b_cs == 1) { k2 = i % SG_K; nn = i / SG_K; } else { nn = i % SG_T; k2 = i / SG_T; }
      const int gn = n0 + nn, gk2 = k0 + k2;
      Bs[k2][nn] = (gn < N && gk2 < K) ? __ldg(b + gn * b_rs + gk2 * b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      if (split) {
        atomicAdd(c + gm * ldc + gn, acc[i][j]);
        continue;
      }
      float v = acc[i][j] + (bias ? __ldg(bias + gn) : 0.f);
      if (accumulate) v += c[gm * ldc + gn];
      if (relu) v = fmaxf(v, 0.f);
      c[gm * ldc + gn] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// Training statistics -> per-(frame, channel) affine + sequential running-stat EMA
// (nn.BatchNorm2d called once per processed frame, models/film_attn_pt_stem.py:211, quirk Q6).
__global__ void bn_finalize_train_kernel(const double* __restrict__ sum, const double* __restrict__ sq,
                                         const int* __restrict__ frame_ct, int n_frames, int C,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         float momentum, float eps, float* __restrict__ scale,
                                         float* __restrict__ shift, float* __restrict__ mean_out,
                                         float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rm = running_mean[c], rv = running_var[c];
  const float g = gamma[c], b = beta[c];
  for (int f = 0; f < n_frames; ++f) {
    const double n = (double)frame_ct[f] * VNQA_IMG_PIX;
    const double mean = sum[(long long)f * C + c] / n;
    double var = sq[(long long)f * C + c] / n - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf((float)var + eps);
    const float sc = g * rstd;
    scale[(long long)f * C + c] = sc;
    shift[(long long)f * C + c] = b - (float)mean * sc;
    mean_out[(long long)f * C + c] = (float)mean;
    rstd_out[(long long)f * C + c] = rstd;
    rm = (1.f - momentum) * rm + momentum * (float)mean;
    rv = (1.f - momentum) * rv + momentum * (float)(var * n / (n - 1.0));
  }
  running_mean[c] = rm;
  running_var[c] = rv;
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

// x[row][c] = y[row][c] * scale[frame(img)][c] + shift[frame(img)][c]   (float4 over channels)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float4* __restrict__ y, float4* __restrict__ x,
                                                       const int* __restrict__ img_frame,
                                                       const float4* __restrict__ scale,
                                                       const float4* __restrict__ shift, long long n_vec, int C4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long row = i / C4;
    const int f = __ldg(img_frame + row / VNQA_IMG_PIX);
    const float4 v = y[i], s = __ldg(scale + (long long)f * C4 + c4), t = __ldg(shift + (long long)f * C4 + c4);
    x[i] = make_float4(tf32_round(fmaf(v.x, s.x, t.x)), tf32_round(fmaf(v.y, s.y, t.y)), tf32_round(fmaf(v.z, s.z, t.z)),
                       tf32_round(fmaf(v.w, s.w, t.w)));
  }
}

// BatchNorm backward, pass 1: per (frame, channel) s1 = sum dx, s2 = sum dx * xhat  (xhat = (y-mean)*rstd).
// One block per (image, 128-channel slab): thread = channel (coalesced 512 B rows), 130 pixels, one atomic pair.
__global__ void __launch_bounds__(128) bn_bwd_reduce_kernel(const float* __restrict__ dx, const float* __restrict__ y,
                                                            const int* __restrict__ img_frame,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, int C,
                                                            float* __restrict__ s1, float* __restrict__ s2) {
  const int img = blockIdx.y;                     // channel slab is the fastest grid index: co-scheduled blocks
  const int c = blockIdx.x * 128 + threadIdx.x;   // sweep whole pixel rows (full DRAM pages)
  if (c >= C) return;
  const int f = img_frame[img];
  const float m = mean[(long long)f * C + c], r = rstd[(long long)f * C + c];
  const float* dp = dx + (long long)img * VNQA_IMG_PIX * C + c;
  const float* yp = y + (long long)img * VNQA_IMG_PIX * C + c;
  float a1 = 0.f, a2 = 0.f;
#pragma unroll 10
  for (int p = 0; p < VNQA_IMG_PIX; ++p) {
    const float d = __ldg(dp + (long long)p * C);
    a1 += d;
    a2 = fmaf(d, (__ldg(yp + (long long)p * C) - m) * r, a2);
  }
  atomicAdd(s1 + (long long)f * C + c, a1);
  atomicAdd(s2 + (long long)f * C + c, a2);
}

// pass 2: dy = gamma*rstd*(dx - s1/n - xhat*s2/n) masked by the ReLU before BN (y > 0); also per-channel
// sum of dy (conv_init bias gradient).  dgamma = sum_f s2, dbeta = sum_f s1 are reduced by the caller kernel below.
__global__ void __launch_bounds__(128) bn_bwd_apply_kernel(const float* __restrict__ dx, const float* __restrict__ y,
                                                           const int* __restrict__ img_frame,
                                                           const int* __restrict__ frame_ct,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ s1, const float* __restrict__ s2,
                                                           int C, float* __restrict__ dy, float* __restrict__ dbias) {
  const int img = blockIdx.y;
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  const int f = img_frame[img];
  const float inv_n = 1.f / ((float)frame_ct[f] * VNQA_IMG_PIX);
  const long long fc = (long long)f * C + c;
  const float m = mean[fc], r = rstd[fc], g = gamma[c] * r;
  const float k1 = s1[fc] * inv_n, k2 = s2[fc] * inv_n;
  const long long base = (long long)img * VNQA_IMG_PIX * C + c;
  float acc = 0.f;
#pragma unroll 10
  for (int p = 0; p < VNQA_IMG_PIX; ++p) {
    const long long i = base + (long long)p * C;
    const float yy = __ldg(y + i);
    const float xhat = (yy - m) * r;
    const float d = yy > 0.f ? tf32_round(g * (__ldg(dx + i) - k1 - xhat * k2)) : 0.f;
    dy[i] = d;
    acc += d;
  }
  atomicAdd(dbias + c, acc);
}

__global__ void bn_bwd_params_kernel(const float* __restrict__ s1, const float* __restrict__ s2, int n_frames, int C,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int f = 0; f < n_frames; ++f) {
    a += s2[(long long)f * C + c];
    b += s1[(long long)f * C + c];
  }
  dgamma[c] += a;
  dbeta[c] += b;
}

// ------------------------------------------------------------------------------------------------ FiLM backward
// Given dOut (gradient of relu(alpha*z+beta)+res wrt the FiLM branch), saved z and the FiLM rows:
//   dF = dOut * [alpha*z+beta > 0]; dz = alpha*dF; dalpha[img][c] = sum_pix dF*z; dbeta[img][c] = sum_pix dF;
//   dbias[c] += sum dz.  Autograd of models/film_attn_pt_stem.py:229-239.  One block per (image, 128-ch slab).
__global__ void __launch_bounds__(256) film_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                                       const float* __restrict__ film,
                                                       const int* __restrict__ film_row, int film_ld, int alpha_off,
                                                       int C, float* __restrict__ dz, float* __restrict__ dfilm,
                                                       float* __restrict__ dbias, int mask_relu) {
  // one block per image; a thread owns 4 consecutive channels (float4), so a warp step reads 512 contiguous bytes and
  // the block sweeps whole 4*C-byte pixel rows: full DRAM pages, 16-byte transactions
  const int img = blockIdx.x;
  const int C4 = C >> 2;
  const long long frow = (long long)film_row[img] * film_ld + alpha_off;
  for (int q = threadIdx.x; q < C4; q += blockDim.x) {
    const float4 al = *reinterpret_cast<const float4*>(film + frow + 4 * q);
    const float4 be = *reinterpret_cast<const float4*>(film + frow + C + 4 * q);
    const float4* zp = reinterpret_cast<const float4*>(z + (long long)img * VNQA_IMG_PIX * C) + q;
    const float4* dp = reinterpret_cast<const float4*>(dout + (long long)img * VNQA_IMG_PIX * C) + q;
    float4* op = reinterpret_cast<float4*>(dz + (long long)img * VNQA_IMG_PIX * C) + q;
    float4 da = make_float4(0.f, 0.f, 0.f, 0.f), db = da, ds = da;
#pragma unroll 5
    for (int p = 0; p < VNQA_IMG_PIX; ++p) {
      const float4 zz = __ldg(zp + (long long)p * C4);
      const float4 dd = __ldg(dp + (long long)p * C4);
      float4 o;
      {
        const float d = fmaf(al.x, zz.x, be.x) > 0.f ? dd.x : 0.f;
        da.x = fmaf(d, zz.x, da.x); db.x += d; o.x = tf32_round(al.x * d); ds.x += o.x;
      }
      {
        const float d = fmaf(al.y, zz.y, be.y) > 0.f ? dd.y : 0.f;
        da.y = fmaf(d, zz.y, da.y); db.y += d; o.y = tf32_round(al.y * d); ds.y += o.y;
      }
      {
        const float d = fmaf(al.z, zz.z, be.z) > 0.f ? dd.z : 0.f;
        da.z = fmaf(d, zz.z, da.z); db.z += d; o.z = tf32_round(al.z * d); ds.z += o.z;
      }
      {
        const float d = fmaf(al.w, zz.w, be.w) > 0.f ? dd.w : 0.f;
        da.w = fmaf(d, zz.w, da.w); db.w += d; o.w = tf32_round(al.w * d); ds.w += o.w;
      }
      op[(long long)p * C4] = o;
    }
    // each (film row, channel) is owned by exactly one image: plain stores.  When the FiLM generator ends in a ReLU
    // (:179, mask_relu) its gradient is masked where the output is 0
#define KEEP(v) (!mask_relu || (v) > 0.f)
    *reinterpret_cast<float4*>(dfilm + frow + 4 * q) =
        make_float4(KEEP(al.x) ? da.x : 0.f, KEEP(al.y) ? da.y : 0.f, KEEP(al.z) ? da.z : 0.f, KEEP(al.w) ? da.w : 0.f);
    *reinterpret_cast<float4*>(dfilm + frow + C + 4 * q) =
        make_float4(KEEP(be.x) ? db.x : 0.f, KEEP(be.y) ? db.y : 0.f, KEEP(be.z) ? db.z : 0.f, KEEP(be.w) ? db.w : 0.f);
#undef KEEP
    atomicAdd(dbias + 4 * q + 0, ds.x);
    atomicAdd(dbias + 4 * q + 1, ds.y);
    atomicAdd(dbias + 4 * q + 2, ds.z);
    atomicAdd(dbias + 4 * q + 3, ds.w);
  }
}

// ------------------------------------------------------------------------------------------------ misc
// out[row][n] = bias[n] + sum_s partial[s][row][n]; rows scattered: out row = dst_row[row] (or row if null)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long rows, int N,
                                     const float* __restrict__ bias, const int* __restrict__ dst_row,
                                     float* __restrict__ out) {
  const long long total = rows * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long row = i / N;
    float a = bias ? __ldg(bias + n) : 0.f;
    for (int s = 0; s < splits; ++s) a += __ldg(partial + (long long)s * total + i);
    const long long orow = dst_row ? dst_row[row] : row;
    out[orow * N + n] = a;
  }
}

// Global temporal max pooling (models/film_global_pooling_pt_stem.py:230-235): y [n_img][130][tail] ->
// pooled[b][c*130 + p] = max_t y[img_of[b][t]][p][c] (values are post-ReLU >= 0, padded rows are 0); argmax image saved.
__global__ void gp_maxpool_kernel(const float* __restrict__ y, const int* __restrict__ img_of, int B, int T, int tail,
                                  float* __restrict__ pooled, int* __restrict__ arg_img) {
  const int total = B * VNQA_IMG_PIX * tail;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % tail;
    const int p = (i / tail) % VNQA_IMG_PIX;
    const int b = i / (tail * VNQA_IMG_PIX);
    float best = 0.f;
    int arg = -1;
    for (int t = 0; t < T; ++t) {
      const int img = img_of[b * T + t];
      if (img < 0) break;
      const float v = __ldg(y + ((long long)img * VNQA_IMG_PIX + p) * tail + c);
      if (arg < 0 || v > best) { best = v; arg = img; }
    }
    const long long o = (long long)b * tail * VNQA_IMG_PIX + (long long)c * VNQA_IMG_PIX + p;
    pooled[o] = arg < 0 ? 0.f : best;
    arg_img[o] = arg;
  }
}

// backward of max pooling + ReLU of the tail conv: dy[n_img][130][tail] (pre-zeroed) gets dpooled at the argmax
// where the pooled value is > 0.
__global__ void gp_maxpool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ pooled,
                                      const int* __restrict__ arg_img, int B, int tail, float* __restrict__ dy) {
  const int total = B * VNQA_IMG_PIX * tail;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int p = i % VNQA_IMG_PIX;
    const int c = (i / VNQA_IMG_PIX) % tail;
    const int img = arg_img[i];
    if (img >= 0 && pooled[i] > 0.f) dy[((long long)img * VNQA_IMG_PIX + p) * tail + c] = dpooled[i];
  }
}

// out[n] (+)= sum_r x[r][n]; grid.x over column blocks of 32, 8 row-lanes per block reduce through smem
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int rows, int n, long long ld,
                                                     float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  float acc = 0.f;
  if (col < n)
    for (int r = rl; r < rows; r += 8) acc += __ldg(x + (long long)r * ld + col);
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && col < n) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    out[col] = accumulate ? out[col] + s : s;
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ src_row, long long rows, int N,
                                   float* __restrict__ out, int round_tf32) {
  const long long total = rows * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    const float v = __ldg(src + (long long)src_row[r] * N + (i % N));
    out[i] = round_tf32 ? tf32_round(v) : v;
  }
}

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_colsum(const float* x, int32_t rows, int32_t n, int64_t ld, float* out, int32_t accumulate,
                           void* stream) {
  VNQA_REQUIRE(x && out && rows > 0 && n > 0, "vnqa_colsum: bad arguments");
  colsum_kernel<<<(n + 31) / 32, 256, 0, as_stream(stream)>>>(x, rows, n, ld, out, accumulate);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_gather_rows(const float* src, const int32_t* src_row, int32_t rows, int32_t N, float* out,
                                int32_t round_tf32, void* stream) {
  const long long total = (long long)rows * N;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  gather_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, src_row, rows, N, out, round_tf32);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_sgemm(const float* a, const float* b, float* c, const float* bias, int32_t M, int32_t N, int32_t K,
                          int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t ldc, int32_t relu,
                          int32_t accumulate, int32_t ksplit, void* stream) {
  VNQA_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0, "vnqa_sgemm: bad arguments");
  if (ksplit < 1) ksplit = 1;
  VNQA_REQUIRE(ksplit == 1 || (!relu && !bias), "vnqa_sgemm: split-K cannot fuse bias/relu");
  int kps = ((K + ksplit - 1) / ksplit + SG_K - 1) / SG_K * SG_K;
  ksplit = (K + kps - 1) / kps;
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T, ksplit);
  sgemm_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, c, bias, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, relu,
                                                    accumulate, kps);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_bn_finalize_train(const double* stat_sum, const double* stat_sq, const int32_t* frame_ct,
                                      int32_t n_frames, int32_t C, const float* gamma, const float* beta,
                                      float* running_mean, float* running_var, float momentum, float eps,
                                      float* scale, float* shift, float* mean, float* rstd, void* stream) {
  VNQA_REQUIRE(stat_sum && stat_sq && frame_ct && scale && shift && mean && rstd, "vnqa_bn_finalize_train: null");
  bn_finalize_train_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(
      stat_sum, stat_sq, frame_ct, n_frames, C, gamma, beta, running_mean, running_var, momentum, eps, scale, shift,
      mean, rstd);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean,
                                   const float* running_var, float eps, int32_t C, float* scale, float* shift,
                                   void* stream) {
  bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(gamma, beta, running_mean, running_var, eps, C,
                                                                        scale, shift);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_bn_apply(const float* y, void* x_, const int32_t* img_frame, const float* scale, const float* shift,
                             int32_t n_img, int32_t C, int32_t act_dtype, void* stream) {
  if (act_dtype == VNQA_ACT_FP16) return bn_apply_h(y, x_, img_frame, scale, shift, n_img, C, as_stream(stream));
  float* x = static_cast<float*>(x_);
  VNQA_REQUIRE(C % 4 == 0, "vnqa_bn_apply: C must be a multiple of 4");
  const long long n_vec = (long long)n_img * VNQA_IMG_PIX * C / 4;
  const int blocks = (int)((n_vec + 255) / 256 < 148 * 8 ? (n_vec + 255) / 256 : 148 * 8);
  bn_apply_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const float4*)y, (float4*)x, img_frame, (const float4*)scale,
                                                        (const float4*)shift, n_vec, C / 4);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_bn_backward(const void* dx_, const float* y, const int32_t* img_frame, const int32_t* frame_ct,
                                int32_t n_frames, const float* mean, const float* rstd, const float* gamma,
                                int32_t n_img, int32_t C, float* s1, float* s2, void* dy_, float* dbias, float* dgamma,
                                float* dbeta, int32_t act_dtype, float grad_unscale, void* stream) {
  VNQA_REQUIRE(dx_ && y && s1 && s2 && dy_ && dbias && dgamma && dbeta, "vnqa_bn_backward: null");
  cudaStream_t st = as_stream(stream);
  if (act_dtype == VNQA_ACT_FP16)
    return bn_backward_h(dx_, y, img_frame, frame_ct, n_frames, mean, rstd, gamma, n_img, C, s1, s2, dy_, dbias, dgamma,
                         dbeta, grad_unscale == 0.f ? 1.f : grad_unscale, st);
  const float* dx = static_cast<const float*>(dx_);
  float* dy = static_cast<float*>(dy_);
  VNQA_CHECK_CUDA(cudaMemsetAsync(s1, 0, (size_t)n_frames * C * 4, st));
  VNQA_CHECK_CUDA(cudaMemsetAsync(s2, 0, (size_t)n_frames * C * 4, st));
  dim3 grid((C + 127) / 128, n_img);
  bn_bwd_reduce_kernel<<<grid, 128, 0, st>>>(dx, y, img_frame, mean, rstd, C, s1, s2);
  bn_bwd_apply_kernel<<<grid, 128, 0, st>>>(dx, y, img_frame, frame_ct, mean, rstd, gamma, s1, s2, C, dy, dbias);
  bn_bwd_params_kernel<<<(C + 127) / 128, 128, 0, st>>>(s1, s2, n_frames, C, dgamma, dbeta);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_film_backward(const void* dout_, const void* z_, const float* film, const int32_t* film_row,
                                  int32_t film_ld, int32_t alpha_off, int32_t n_img, int32_t C, void* dz_, float* dfilm,
                                  float* dbias, int32_t mask_relu, int32_t act_dtype, float grad_unscale, void* stream) {
  VNQA_REQUIRE(dout_ && z_ && film && film_row && dz_ && dfilm && dbias, "vnqa_film_backward: null");
  if (act_dtype == VNQA_ACT_FP16)
    return film_backward_h(dout_, z_, film, film_row, film_ld, alpha_off, n_img, C, dz_, dfilm, dbias, mask_relu,
                           grad_unscale == 0.f ? 1.f : grad_unscale, as_stream(stream));
  const float* dout = static_cast<const float*>(dout_);
  const float* z = static_cast<const float*>(z_);
  float* dz = static_cast<float*>(dz_);
  VNQA_REQUIRE(C % 4 == 0 && film_ld % 4 == 0 && alpha_off % 4 == 0, "vnqa_film_backward: C/film_ld/alpha_off must be multiples of 4");
  const int threads = C / 4 >= 256 ? 256 : ((C / 4 + 31) / 32) * 32;
  film_bwd_kernel<<<n_img, threads, 0, as_stream(stream)>>>(dout, z, film, film_row, film_ld, alpha_off, C, dz, dfilm,
                                                            dbias, mask_relu);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_splitk_reduce(const float* partial, int32_t splits, int32_t rows, int32_t N, const float* bias,
                                  const int32_t* dst_row, float* out, void* stream) {
  const long long total = (long long)rows * N;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  splitk_reduce_kernel<<<blocks, 256, 0, as_stream(stream)>>>(partial, splits, rows, N, bias, dst_row, out);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_gp_maxpool(const float* y, const int32_t* img_of, int32_t B, int32_t T, int32_t tail, float* pooled,
                               int32_t* arg_img, void* stream) {
  const int total = B * VNQA_IMG_PIX * tail;
  gp_maxpool_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(y, img_of, B, T, tail, pooled, arg_img);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_gp_maxpool_backward(const float* dpooled, const float* pooled, const int32_t* arg_img, int32_t B,
                                        int32_t tail, float* dy, void* stream) {
  const int total = B * VNQA_IMG_PIX * tail;
  gp_maxpool_bwd_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(dpooled, pooled, arg_img, B, tail, dy);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
