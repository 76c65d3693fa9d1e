// Global-norm gradient clipping fused into Adam over FLAT parameter / gradient / moment buffers
// (replaces clip_grad_norm(model.parameters(), 1.0) + optimizer.step() of
// eval/q_and_v_eval.py:137-139 with two HBM-bound launches instead of ~40 foreach launches).
//
//   pass 1  vnqa_sumsq_partials : per-block sums of g^2 in fp64, one partial per block (no atomics: deterministic)
//   pass 2  vnqa_clip_adam      : every block re-reduces the partials in the same fixed order, forms
//                                 coef = min(1, max_norm / (||g|| + 1e-6))            (torch.nn.utils.clip_grad_norm_)
//                                 and applies torch.optim.Adam's update (no amsgrad, no weight decay):
//                                 m += (1-b1)(g-m);  v = b2 v + (1-b2) g^2;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// Algorithmic traffic: pass 1 reads 4 B/element, pass 2 reads 16 B and writes 12 B per element.
#include <cmath>

#include "common.h"

namespace vnqa {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) sumsq_partials_kernel(const float4* __restrict__ g, long long n4,
                                                                     double* __restrict__ partials) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  double acc = 0.0;
  int since = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g + i);
    a0 = fmaf(v.x, v.x, a0); a1 = fmaf(v.y, v.y, a1); a2 = fmaf(v.z, v.z, a2); a3 = fmaf(v.w, v.w, a3);
    if (++since == 64) {   // bound the fp32 chain length
      acc += (double)a0 + (double)a1 + (double)a2 + (double)a3;
      a0 = a1 = a2 = a3 = 0.f; since = 0;
    }
  }
  acc += (double)a0 + (double)a1 + (double)a2 + (double)a3;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double red[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < OPT_THREADS / 32; ++i) s += red[i];
    partials[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) clip_adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                                float4* __restrict__ m, float4* __restrict__ v,
                                                                long long n4, const double* __restrict__ partials,
                                                                int n_partials, float grad_scale, float max_norm,
                                                                float step_size, float beta2, float w1, float w2,
                                                                float eps, float bias_correction2_sqrt,
                                                                float* __restrict__ norm_out,
                                                                const long long* __restrict__ step_dev, double lr,
                                                                double beta1_d, double beta2_d) {
  // every block sums the same partials in the same order -> identical coef everywhere, run to run
  __shared__ double red[OPT_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) s += partials[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (step_dev != nullptr) {
    // capturable mode (CUDA graphs): the step count lives on the device, so the bias corrections are formed here -
    // in double, exactly like the host path - by one thread per block
    __shared__ float sc[2];
    if (threadIdx.x == 0) {
      const double t = (double)*step_dev;
      sc[0] = (float)(lr / (1.0 - pow(beta1_d, t)));
      sc[1] = (float)sqrt(1.0 - pow(beta2_d, t));
    }
    __syncthreads();
    step_size = sc[0];
    bias_correction2_sqrt = sc[1];
  }
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float total_norm = (float)sqrt(red[0]) * grad_scale;
    float c = max_norm / (total_norm + 1e-6f);
    coef *= c < 1.f ? c : 1.f;
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out != nullptr) *norm_out = total_norm;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 gi = __ldg(g + i);
    float4 mi = m[i], vi = v[i], pi = p[i];
#define VNQA_ADAM1(c)                                                     \
    {                                                                     \
      const float gg = gi.c * coef;                                       \
      mi.c = mi.c + w1 * (gg - mi.c);                                     \
      vi.c = fmaf(w2 * gg, gg, beta2 * vi.c);                             \
      const float denom = sqrtf(vi.c) / bias_correction2_sqrt + eps;      \
      pi.c = pi.c - step_size * (mi.c / denom);                           \
    }
    VNQA_ADAM1(x) VNQA_ADAM1(y) VNQA_ADAM1(z) VNQA_ADAM1(w)
#undef VNQA_ADAM1
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}

__global__ void adam_step_inc_kernel(long long* step_dev) { *step_dev += 1; }

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_opt_num_partials(void) {
  const int sms = vnqa_sm_count();
  return sms > 0 ? sms * 4 : -2;
}

extern "C" int vnqa_sumsq_partials(const float* g, int64_t n, double* partials, void* stream) {
  VNQA_REQUIRE(g && partials && n > 0 && n % 4 == 0, "vnqa_sumsq_partials: need a non-empty buffer of 4k floats");
  VNQA_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "vnqa_sumsq_partials: buffer must be 16-byte aligned");
  const int blocks = vnqa_opt_num_partials();
  if (blocks <= 0) return -2;
  sumsq_partials_kernel<<<blocks, OPT_THREADS, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(g), n / 4,
                                                                        partials);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int clip_adam_launch(float* p, const float* g, float* m, float* v, int64_t n, const double* partials,
                            int32_t n_partials, float grad_scale, float max_norm, double lr, double beta1, double beta2,
                            double eps, int64_t step, long long* step_dev, float* norm_out, void* stream) {
  // scalar coefficients exactly as torch.optim.Adam forms them: in double (on the host, or by the kernel from the
  // device-resident step count), rounded to fp32 once
  VNQA_REQUIRE(step_dev != nullptr || step >= 1, "vnqa_clip_adam: step counts from 1");
  const double tt = step_dev ? 1.0 : (double)step;
  const double bc1 = 1.0 - pow(beta1, tt), bc2 = 1.0 - pow(beta2, tt);
  const float step_size = (float)(lr / bc1), bias_correction2_sqrt = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2);
  VNQA_REQUIRE(p && g && m && v && n > 0 && n % 4 == 0, "vnqa_clip_adam: need non-empty buffers of 4k floats");
  VNQA_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "vnqa_clip_adam: buffers must be 16-byte aligned");
  VNQA_REQUIRE(max_norm <= 0.f || (partials && n_partials > 0), "vnqa_clip_adam: clipping needs the sum-of-squares partials");
  const int sms = vnqa_sm_count();
  if (sms <= 0) return -2;
  clip_adam_kernel<<<sms * 8, OPT_THREADS, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), n / 4, partials, n_partials, grad_scale, max_norm, step_size, (float)beta2, w1, w2,
      (float)eps, bias_correction2_sqrt, norm_out, step_dev, lr, beta1, beta2);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const double* partials,
                              int32_t n_partials, float grad_scale, float max_norm, double lr, double beta1,
                              double beta2, double eps, int64_t step, float* norm_out, void* stream) {
  return clip_adam_launch(p, g, m, v, n, partials, n_partials, grad_scale, max_norm, lr, beta1, beta2, eps, step,
                          nullptr, norm_out, stream);
}

extern "C" int vnqa_clip_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const double* partials,
                                  int32_t n_partials, float grad_scale, float max_norm, double lr, double beta1,
                                  double beta2, double eps, int64_t* step_dev, float* norm_out, void* stream) {
  VNQA_REQUIRE(step_dev != nullptr, "vnqa_clip_adam_dev: null step counter");
  adam_step_inc_kernel<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<long long*>(step_dev));
  VNQA_CHECK_CUDA(cudaGetLastError());
  return clip_adam_launch(p, g, m, v, n, partials, n_partials, grad_scale, max_norm, lr, beta1, beta2, eps, 0,
                          reinterpret_cast<long long*>(step_dev), norm_out, stream);
}
