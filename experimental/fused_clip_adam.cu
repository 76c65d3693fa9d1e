// EXPERIMENTAL (not built, not validated on hardware yet) - SURVEY.md §8(f) rank 1.
// clip_grad_norm_(params, max_norm) + Adam.step() of eval/q_and_v_eval.py:137-139,333 as two multi-tensor launches:
//   1. grad_sqnorm_kernel: sum of squares of every gradient element -> one fp64 scalar (HBM-bound, 283 MB read)
//   2. clip_adam_kernel:   g *= min(1, max_norm / (||g|| + 1e-6)); standard (non-amsgrad, no weight decay) Adam update
//                          with bias correction, exactly torch.optim.Adam's formula (HBM-bound: reads g, p, m, v;
//                          writes p, m, v)
// Tensors are addressed through a chunk table so that one launch covers all ~30 parameter tensors.
#include <cstdint>
#include <cuda_runtime.h>

namespace vnqa_exp {

constexpr int CHUNK = 1 << 16;   // elements per block

struct ChunkTable {
  const int* tensor_of_chunk;     // [n_chunks]
  const long long* chunk_start;   // [n_chunks] element offset inside its tensor
  float* const* param;            // [n_tensors]
  const float* const* grad;
  float* const* exp_avg;
  float* const* exp_avg_sq;
  const long long* numel;         // [n_tensors]
};

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const ChunkTable t, double* __restrict__ norm2) {
  const int c = blockIdx.x;
  const int ti = t.tensor_of_chunk[c];
  const long long start = t.chunk_start[c];
  const long long n = min((long long)CHUNK, t.numel[ti] - start);
  const float* g = t.grad[ti] + start;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += (double)red[i];
    atomicAdd(norm2, s);
  }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(const ChunkTable t, const double* __restrict__ norm2,
                                                        float max_norm, float lr, float beta1, float beta2, float eps,
                                                        float bias_correction1, float bias_correction2_sqrt) {
  const int c = blockIdx.x;
  const int ti = t.tensor_of_chunk[c];
  const long long start = t.chunk_start[c];
  const long long n = min((long long)CHUNK, t.numel[ti] - start);
  const float total_norm = (float)sqrt(*norm2);
  float coef = max_norm / (total_norm + 1e-6f);          // torch.nn.utils.clip_grad_norm_
  coef = coef < 1.f ? coef : 1.f;
  const float* g = t.grad[ti] + start;
  float* p = t.param[ti] + start;
  float* m = t.exp_avg[ti] + start;
  float* v = t.exp_avg_sq[ti] + start;
  const float step_size = lr / bias_correction1;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias_correction2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

}  // namespace vnqa_exp

extern "C" int vnqa_exp_fused_clip_adam(const int* tensor_of_chunk, const long long* chunk_start, float* const* param,
                                        const float* const* grad, float* const* exp_avg, float* const* exp_avg_sq,
                                        const long long* numel, int n_chunks, double* norm2_scratch, float max_norm,
                                        float lr, float beta1, float beta2, float eps, float bias_correction1,
                                        float bias_correction2_sqrt, void* stream) {
  using namespace vnqa_exp;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ChunkTable t{tensor_of_chunk, chunk_start, param, grad, exp_avg, exp_avg_sq, numel};
  if (cudaMemsetAsync(norm2_scratch, 0, sizeof(double), st) != cudaSuccess) return -2;
  grad_sqnorm_kernel<<<n_chunks, 256, 0, st>>>(t, norm2_scratch);
  clip_adam_kernel<<<n_chunks, 256, 0, st>>>(t, norm2_scratch, max_norm, lr, beta1, beta2, eps, bias_correction1,
                                             bias_correction2_sqrt);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
