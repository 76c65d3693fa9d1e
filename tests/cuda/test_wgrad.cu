// Standalone GPU check + micro-benchmark of vnqa_conv_wgrad against a naive double-precision CUDA reference.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/vnqa_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define CKV(x) do { int r_ = (x); if (r_ != 0) { printf("vnqa error %d: %s (%s:%d)\n", r_, vnqa_last_error(), __FILE__, __LINE__); exit(3); } } while (0)

__global__ void fill_kernel(float* x, long long n, uint32_t seed, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    float v = ((h >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v * scale));
    x[i] = __uint_as_float(r);
  }
}

// dw[co][ci][tap] for sampled (co, ci, tap)
__global__ void ref_kernel(const float* __restrict__ dy, const float* __restrict__ x, double* __restrict__ out,
                           const int* __restrict__ samples, int ns, int cout, int cin, int taps, int n_img, int flat) {
  const int s = blockIdx.x;
  if (s >= ns) return;
  const int co = samples[3 * s], ci = samples[3 * s + 1], tap = samples[3 * s + 2];
  int dyy = 0, dxx = 0;
  if (taps == 9) { dyy = tap / 3 - 1; dxx = tap % 3 - 1; }
  const long long rows = flat ? n_img : (long long)n_img * 130;
  double acc = 0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    long long xr = r;
    if (!flat) {
      const int img = (int)(r / 130), pix = (int)(r % 130), h = pix / 13 + dyy, w = pix % 13 + dxx;
      if (h < 0 || h >= 10 || w < 0 || w >= 13) continue;
      xr = (long long)img * 130 + h * 13 + w;
    }
    acc += (double)dy[r * cout + co] * (double)x[xr * cin + ci];
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[s] = red[0];
}

struct Case { const char* name; int taps, cin, cout, n_img, flat, iters; };

static int run_case(const Case& c) {
  const long long rows = c.flat ? c.n_img : (long long)c.n_img * 130;
  float *dy, *x, *dw, *ws;
  CK(cudaMalloc(&dy, rows * c.cout * 4));
  CK(cudaMalloc(&x, rows * c.cin * 4));
  const size_t wn = (size_t)c.cout * c.cin * c.taps;
  CK(cudaMalloc(&dw, wn * 4));
  const int64_t ws_bytes = vnqa_conv_wgrad_ws_bytes(c.cout, c.cin, c.taps, nullptr);
  CK(cudaMalloc(&ws, ws_bytes));
  CK(cudaMemset(ws, 0, ws_bytes));   // only the leading turn counters need it; the kernel leaves them zero
  fill_kernel<<<148 * 8, 256>>>(dy, rows * c.cout, 17u, 1.0f);
  fill_kernel<<<148 * 8, 256>>>(x, rows * c.cin, 99u, 1.0f);
  CK(cudaMemset(dw, 0xFF, wn * 4));   // garbage: the kernel overwrites dw
  CK(cudaDeviceSynchronize());
  fprintf(stderr, "[%s] inputs ready\n", c.name);
  CKV(vnqa_conv_wgrad(dy, x, dw, ws, c.n_img, c.cout, c.cin, c.taps, c.flat, VNQA_ACT_TF32, 1.f, nullptr));
  CK(cudaDeviceSynchronize());
  fprintf(stderr, "[%s] wgrad done\n", c.name);
  const int ns = 512;
  std::vector<int> hs(3 * ns);
  uint32_t seed = 777;
  for (int i = 0; i < ns; ++i) {
    seed = seed * 1664525u + 1013904223u; hs[3 * i] = (seed >> 8) % c.cout;
    seed = seed * 1664525u + 1013904223u; hs[3 * i + 1] = (seed >> 8) % c.cin;
    seed = seed * 1664525u + 1013904223u; hs[3 * i + 2] = (seed >> 8) % c.taps;
  }
  // always include the corners
  hs[0] = 0; hs[1] = 0; hs[2] = 0;
  hs[3] = c.cout - 1; hs[4] = c.cin - 1; hs[5] = c.taps - 1;
  int* ds; double* dref;
  CK(cudaMalloc(&ds, hs.size() * 4));
  CK(cudaMemcpy(ds, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dref, ns * 8));
  ref_kernel<<<ns, 256>>>(dy, x, dref, ds, ns, c.cout, c.cin, c.taps, c.n_img, c.flat);
  CK(cudaDeviceSynchronize());
  if (getenv("VNQA_TEST_VERBOSE")) {
    std::vector<float> h0(wn);
    CK(cudaMemcpy(h0.data(), dw, wn * 4, cudaMemcpyDeviceToHost));
    size_t nz = 0; double mxv = 0;
    for (size_t i = 0; i < wn; ++i) { if (h0[i] != 0.f) ++nz; mxv = fmax(mxv, fabs(h0[i])); }
    printf("   dw: %zu of %zu non-zero, max |dw| %.4f\n", nz, wn, mxv);
  }
  std::vector<double> href(ns);
  std::vector<float> hw(wn);
  CK(cudaMemcpy(href.data(), dref, ns * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hw.data(), dw, wn * 4, cudaMemcpyDeviceToHost));
  double err = 0, mx = 0;
  for (int i = 0; i < ns; ++i) {
    const double got = hw[((size_t)hs[3 * i] * c.cin + hs[3 * i + 1]) * c.taps + hs[3 * i + 2]];
    if (i < 4 && getenv("VNQA_TEST_VERBOSE")) printf("   sample co %d ci %d tap %d: got %.6f ref %.6f\n", hs[3*i], hs[3*i+1], hs[3*i+2], got, href[i]);
    err = fmax(err, fabs(got - href[i]));
    mx = fmax(mx, fabs(href[i]));
  }
  const int fail = !(err / mx < (c.iters > 0 ? 2e-3 : 2e-4));   // long fp32 accumulation chains in the perf cases
  printf("%-26s taps %d cin %6d cout %5d n_img %5d  normalised max err %.3e", c.name, c.taps, c.cin, c.cout, c.n_img, err / mx);
  if (c.iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CKV(vnqa_conv_wgrad(dy, x, dw, ws, c.n_img, c.cout, c.cin, c.taps, c.flat, VNQA_ACT_TF32, 1.f, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.iters; ++i) CKV(vnqa_conv_wgrad(dy, x, dw, ws, c.n_img, c.cout, c.cin, c.taps, c.flat, VNQA_ACT_TF32, 1.f, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= c.iters;
    printf("  | %.3f ms  %.1f TFLOP/s", ms, 2.0 * rows * c.cout * (double)c.cin * c.taps / ms * 1e-9);
  }
  printf("  %s\n", fail ? "FAIL" : "ok");
  fflush(stdout);
  cudaFree(dy); cudaFree(x); cudaFree(dw); cudaFree(ws); cudaFree(ds); cudaFree(dref);
  return fail;
}

int main(int argc, char** argv) {
  bool perf = false;
  const char* only = nullptr;   // --case <substring>: run only the matching cases (used under ncu)
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--perf")) perf = true;
    if (!strcmp(argv[i], "--case") && i + 1 < argc) only = argv[++i];
  }
  CKV(vnqa_device_check());
  std::vector<Case> cases = {
      {"wgrad1x1_small", 1, 128, 128, 8, 0, 0},
      {"wgrad3x3_small", 9, 128, 128, 8, 0, 0},
      {"wgrad3x3_ragged", 9, 96, 160, 11, 0, 0},
      {"wgrad3x3_narrow", 9, 32, 32, 5, 0, 0},
      {"wgrad_flat", 1, 4160, 128, 300, 1, 0},
      {"wgrad_flat_small", 1, 64, 32, 77, 1, 0},
  };
  if (perf) {
    cases.push_back({"P wgrad conv_init", 9, 512, 1024, 2240, 0, 3});
    cases.push_back({"P wgrad 1x1", 1, 1024, 1024, 2240, 0, 5});
    cases.push_back({"P wgrad 3x3", 9, 1024, 1024, 2240, 0, 3});
    cases.push_back({"P wgrad fc_embed", 1, 133120, 128, 2240, 1, 5});
  }
  int fails = 0;
  for (const auto& c : cases)
    if (!only || strstr(c.name, only)) fails += run_case(c);
  printf("%s (%d failing cases)\n", fails ? "FAILED" : "ALL OK", fails);
  return fails ? 1 : 0;
}
