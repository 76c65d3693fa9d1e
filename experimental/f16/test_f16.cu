// Standalone GPU check + micro-benchmark of the fp16 candidate kernels (vnqa_conv_gemm_h, vnqa_conv_wgrad_h) against
// naive CUDA references with double accumulation over the SAME fp16 operands (so products are exact and only the
// accumulation order / the fp16 rounding of the stored result differ).  NOT YET RUN ON HARDWARE.
//   make -C experimental/f16 && build/test_f16 [--perf] [--case <substring>] [--wgrad-only | --conv-only]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "vnqa_f16.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)
#define CKV(x)                                                          \
  do {                                                                  \
    int r_ = (x);                                                       \
    if (r_ != 0) {                                                      \
      printf("vnqa error %d: %s (%s:%d)\n", r_, vnqa_last_error(), __FILE__, __LINE__); \
      exit(3);                                                          \
    }                                                                   \
  } while (0)

// ---------------------------------------------------------------------------------------------- helpers
__global__ void fill_half_kernel(__half* p, long long n, uint32_t seed, float scale, int relu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    float v = ((h & 0xFFFF) / 32768.0f - 1.0f) * scale;
    if (relu) v = fmaxf(v, 0.f);
    p[i] = __float2half_rn(v);
  }
}
__global__ void fill_float_kernel(float* p, long long n, uint32_t seed, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = ((h & 0xFFFF) / 32768.0f - 1.0f) * scale;
  }
}

// raw accumulator of the convolution / flat GEMM at sampled (row, co); w is the PyTorch-layout fp32 weight, rounded to
// fp16 on the fly exactly as vnqa_pack_conv_weight_h does
__global__ void conv_ref_kernel(const float* __restrict__ w, const __half* __restrict__ x, double* __restrict__ out,
                                const int* __restrict__ samples, long long n_samples, int cout, int cin, int taps,
                                int flat, int x_ld) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n_samples;
       s += (long long)gridDim.x * blockDim.x) {
    long long row;
    int co;
    if (samples) { row = samples[2 * s]; co = samples[2 * s + 1]; }
    else { row = s / cout; co = (int)(s % cout); }
    double acc = 0.0;
    if (flat) {
      for (int k = 0; k < cin; ++k)
        acc += (double)__half2float(__float2half_rn(w[(long long)co * cin + k])) * (double)__half2float(x[row * x_ld + k]);
    } else {
      const int img = (int)(row / 130), pix = (int)(row % 130), h = pix / 13, wq = pix % 13;
      for (int tap = 0; tap < taps; ++tap) {
        int dy = 0, dx = 0;
        if (taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        const int hh = h + dy, ww = wq + dx;
        if (hh < 0 || hh >= 10 || ww < 0 || ww >= 13) continue;
        const __half* xp = x + ((long long)img * 130 + hh * 13 + ww) * cin;
        for (int k = 0; k < cin; ++k)
          acc += (double)__half2float(__float2half_rn(w[((long long)co * cin + k) * taps + tap])) *
                 (double)__half2float(xp[k]);
      }
    }
    out[s] = acc;
  }
}

__global__ void wgrad_ref_kernel(const __half* __restrict__ dy, const __half* __restrict__ x, double* __restrict__ out,
                                 const int* __restrict__ samples, int ns, int cout, int cin, int taps, int n_img,
                                 int flat) {
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
    acc += (double)__half2float(dy[r * cout + co]) * (double)__half2float(x[xr * cin + ci]);
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[s] = red[0];
}

static double half_round_host(double v) { return (double)__half2float(__float2half_rn((float)v)); }
static uint32_t g_seed = 12345;
static uint32_t next_u32() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 4; }

template <class T>
static std::vector<T> from_dev(const void* d, size_t n) {
  std::vector<T> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}

// ---------------------------------------------------------------------------------------------- conv cases
struct ConvCase {
  const char* name;
  int taps, cin, cout, n_img, flat, splits, epi;
  long long samples;  // 0 = check everything
  int iters;          // >0: time it
  int out_f32 = 0;    // BIAS_RELU / PLAIN: fp32 output
};

static int run_conv(const ConvCase& c) {
  const int rows_pad = (c.cout + 127) / 128 * 128;
  const long long rows = c.flat ? c.n_img : (long long)c.n_img * 130;
  const int K = c.flat ? c.cin * c.splits : c.cin;
  const int T = 4;  // frames for STATS
  const int film_ld = 2 * c.cout + 16;
  const bool f32out = c.epi == VNQA_EPI_BIAS_RELU_STATS || c.epi == VNQA_EPI_PARTIAL || c.out_f32;
  const size_t out_elems = (size_t)rows * c.cout * (c.epi == VNQA_EPI_PARTIAL ? c.splits : 1);

  float *dw, *dbias, *dscale, *dshift, *dfilm, *ddb;
  __half *dx, *dres, *dadd, *dwp, *dz;
  void* dout;
  double *dsum, *dsq;
  int *dfilm_row, *dframe;
  CK(cudaMalloc(&dw, (size_t)c.cout * K * c.taps * 4));
  CK(cudaMalloc(&dx, (size_t)rows * K * 2));
  CK(cudaMalloc(&dbias, c.cout * 4)); CK(cudaMalloc(&dscale, c.cout * 4)); CK(cudaMalloc(&dshift, c.cout * 4));
  CK(cudaMalloc(&dres, (size_t)rows * c.cout * 2)); CK(cudaMalloc(&dadd, (size_t)rows * c.cout * 2));
  CK(cudaMalloc(&dz, (size_t)rows * c.cout * 2));
  CK(cudaMalloc(&dfilm, (size_t)c.n_img * film_ld * 4));
  CK(cudaMalloc(&dwp, (size_t)c.taps * rows_pad * K * 2));
  CK(cudaMalloc(&dout, out_elems * (f32out ? 4 : 2)));
  CK(cudaMalloc(&ddb, c.cout * 4));
  CK(cudaMalloc(&dsum, (size_t)T * c.cout * 8)); CK(cudaMalloc(&dsq, (size_t)T * c.cout * 8));
  fill_float_kernel<<<592, 256>>>(dw, (long long)c.cout * K * c.taps, 11u, 0.05f);
  fill_half_kernel<<<592, 256>>>(dx, rows * K, 23u, 1.0f, 1);
  fill_float_kernel<<<8, 256>>>(dbias, c.cout, 31u, 0.5f);
  fill_float_kernel<<<8, 256>>>(dscale, c.cout, 37u, 1.0f);
  fill_float_kernel<<<8, 256>>>(dshift, c.cout, 41u, 1.0f);
  fill_half_kernel<<<592, 256>>>(dres, rows * c.cout, 43u, 1.0f, 0);
  fill_half_kernel<<<592, 256>>>(dadd, rows * c.cout, 47u, 1.0f, 0);
  fill_float_kernel<<<592, 256>>>(dfilm, (long long)c.n_img * film_ld, 53u, 1.0f);
  std::vector<int> hfilm_row(c.n_img), hframe(c.n_img);
  for (int i = 0; i < c.n_img; ++i) {
    hfilm_row[i] = (i * 7 + 3) % c.n_img;
    hframe[i] = (int)((long long)i * T / c.n_img);
  }
  CK(cudaMalloc(&dfilm_row, c.n_img * 4)); CK(cudaMalloc(&dframe, c.n_img * 4));
  CK(cudaMemcpy(dfilm_row, hfilm_row.data(), c.n_img * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dframe, hframe.data(), c.n_img * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, out_elems * (f32out ? 4 : 2)));
  CK(cudaMemset(ddb, 0, c.cout * 4));
  CK(cudaMemset(dsum, 0, (size_t)T * c.cout * 8)); CK(cudaMemset(dsq, 0, (size_t)T * c.cout * 8));
  CKV(vnqa_pack_conv_weight_h(dw, dwp, c.cout, K, c.taps, rows_pad, 0, nullptr));
  CK(cudaDeviceSynchronize());

  vnqa_conv_args_h a;
  memset(&a, 0, sizeof(a));
  a.w_packed = dwp; a.x = dx; a.taps = c.taps; a.cin = c.cin; a.cout = c.cout; a.rows_pad = rows_pad;
  a.n_img = c.n_img; a.flat = c.flat; a.x_ld = K; a.splits = c.splits; a.epi = c.epi; a.out = dout;
  a.bias = dbias; a.scale = dscale; a.shift = dshift; a.img_frame = dframe; a.stat_sum = dsum; a.stat_sq = dsq;
  a.film = dfilm; a.film_row = dfilm_row; a.film_ld = film_ld; a.alpha_off = 8; a.res = dres; a.z_out = dz;
  a.add = dadd; a.dbias = ddb; a.grad_unscale = 0.25f; a.out_f32 = c.out_f32;
  CKV(vnqa_conv_gemm_h(&a, nullptr));
  CK(cudaDeviceSynchronize());

  long long ns = c.samples > 0 ? c.samples : rows * c.cout;
  std::vector<int> hs;
  int* dsamp = nullptr;
  if (c.samples > 0) {
    hs.resize(2 * ns);
    for (long long i = 0; i < ns; ++i) { hs[2 * i] = (int)(next_u32() % rows); hs[2 * i + 1] = (int)(next_u32() % c.cout); }
    CK(cudaMalloc(&dsamp, hs.size() * 4));
    CK(cudaMemcpy(dsamp, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
  }
  double* dref;
  CK(cudaMalloc(&dref, ns * 8));
  conv_ref_kernel<<<592, 256>>>(dw, dx, dref, dsamp, ns, c.cout, K, c.taps, c.flat, K);
  CK(cudaDeviceSynchronize());
  auto href = from_dev<double>(dref, ns);
  auto hbias = from_dev<float>(dbias, c.cout), hscale = from_dev<float>(dscale, c.cout), hshift = from_dev<float>(dshift, c.cout);
  auto hfilm = from_dev<float>(dfilm, (size_t)c.n_img * film_ld);
  auto hres = from_dev<__half>(dres, (size_t)rows * c.cout), hadd = from_dev<__half>(dadd, (size_t)rows * c.cout);
  auto hz = from_dev<__half>(dz, (size_t)rows * c.cout);
  auto hdb = from_dev<float>(ddb, c.cout);
  auto hsum = from_dev<double>(dsum, (size_t)T * c.cout), hsq = from_dev<double>(dsq, (size_t)T * c.cout);
  std::vector<__half> hout16;
  std::vector<float> hout32;
  if (f32out) hout32 = from_dev<float>(dout, out_elems); else hout16 = from_dev<__half>(dout, out_elems);

  double err = 0, maxref = 0;
  std::vector<double> ssum((size_t)T * c.cout, 0.0), ssq((size_t)T * c.cout, 0.0), sdb(c.cout, 0.0);
  int fail = 0;
  for (long long s = 0; s < ns; ++s) {
    const long long row = c.samples > 0 ? hs[2 * s] : s / c.cout;
    const int co = c.samples > 0 ? hs[2 * s + 1] : (int)(s % c.cout);
    const long long idx = row * c.cout + co;
    const int img = c.flat ? 0 : (int)(row / 130);
    const double acc = href[s];
    double expect = 0;
    switch (c.epi) {
      case VNQA_EPI_BIAS_RELU: expect = fmax(acc + hbias[co], 0.0); break;
      case VNQA_EPI_BIAS_RELU_AFFINE: expect = fmax(acc + hbias[co], 0.0) * hscale[co] + hshift[co]; break;
      case VNQA_EPI_BIAS_RELU_STATS:
        expect = fmax(acc + hbias[co], 0.0);
        if (c.samples == 0) {
          ssum[(size_t)hframe[img] * c.cout + co] += expect;
          ssq[(size_t)hframe[img] * c.cout + co] += expect * expect;
        }
        break;
      case VNQA_EPI_FILM: {
        const float* fr = &hfilm[(size_t)hfilm_row[img] * film_ld + 8];
        const double z = acc + hbias[co];
        expect = fmax(fr[co] * z + fr[c.cout + co], 0.0) + __half2float(hres[idx]);
        const double ez = fabs(__half2float(hz[idx]) - z);
        if (ez > 1.1e-3 * (1 + fabs(z)) && !fail) {
          printf("  z_out mismatch at row %lld co %d: %f vs %f\n", row, co, __half2float(hz[idx]), z);
          fail = 1;
        }
        break;
      }
      case VNQA_EPI_PLAIN: expect = acc; break;
      case VNQA_EPI_DGRAD_RES:
        expect = __half2float(hres[idx]) > 0.f ? half_round_host(acc + __half2float(hadd[idx])) : 0.0;
        if (c.samples == 0) sdb[co] += expect;
        break;
      case VNQA_EPI_PARTIAL: expect = acc; break;
    }
    if (!f32out) expect = half_round_host(expect);
    double got;
    if (c.epi == VNQA_EPI_PARTIAL) {
      got = 0;
      for (int sp = 0; sp < c.splits; ++sp) got += hout32[(size_t)sp * rows * c.cout + idx];
    } else {
      got = f32out ? (double)hout32[idx] : (double)__half2float(hout16[idx]);
    }
    err = fmax(err, fabs(got - expect));
    maxref = fmax(maxref, fabs(expect));
  }
  const double n1 = err / maxref;
  printf("%-30s taps %d cin %5d cout %5d n_img %5d  normalised max err vs fp16-operand reference %.3e", c.name, c.taps,
         K, c.cout, c.n_img, n1);
  // fp16 outputs may differ by one fp16 ulp (2^-10 relative) where the fp32 sums differ in the last bits
  if (!(n1 < (f32out ? 1e-4 : 1.1e-3))) fail = 1;
  if (c.epi == VNQA_EPI_BIAS_RELU_STATS && c.samples == 0) {
    double es = 0, ms = 0;
    for (size_t i = 0; i < ssum.size(); ++i) {
      es = fmax(es, fabs(ssum[i] - hsum[i]));
      es = fmax(es, fabs(ssq[i] - hsq[i]));
      ms = fmax(ms, fabs(ssq[i]));
    }
    printf("  stats err %.3e", es / ms);
    if (!(es / ms < 1e-3)) fail = 1;
  }
  if (c.epi == VNQA_EPI_DGRAD_RES && c.samples == 0) {
    double es = 0, ms = 0;
    for (int i = 0; i < c.cout; ++i) {
      es = fmax(es, fabs(0.25 * sdb[i] - hdb[i]));   // grad_unscale = 0.25
      ms = fmax(ms, fabs(0.25 * sdb[i]));
    }
    printf("  dbias err %.3e", es / ms);
    if (!(es / ms < 2e-3)) fail = 1;
  }
  if (c.iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) CKV(vnqa_conv_gemm_h(&a, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.iters; ++i) CKV(vnqa_conv_gemm_h(&a, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= c.iters;
    printf("  | %.3f ms  %.1f TFLOP/s", ms, 2.0 * rows * c.cout * (double)K * c.taps / ms * 1e-9);
  }
  printf("  %s\n", fail ? "FAIL" : "ok");
  fflush(stdout);
  cudaFree(dw); cudaFree(dx); cudaFree(dbias); cudaFree(dscale); cudaFree(dshift); cudaFree(dres); cudaFree(dadd);
  cudaFree(dz); cudaFree(dfilm); cudaFree(dwp); cudaFree(dout); cudaFree(ddb); cudaFree(dsum); cudaFree(dsq);
  cudaFree(dfilm_row); cudaFree(dframe); cudaFree(dref);
  if (dsamp) cudaFree(dsamp);
  return fail;
}

// ---------------------------------------------------------------------------------------------- wgrad cases
struct WgradCase { const char* name; int taps, cin, cout, n_img, flat, iters; };

static int run_wgrad(const WgradCase& c) {
  const long long rows = c.flat ? c.n_img : (long long)c.n_img * 130;
  __half *dy, *x;
  float *dw, *ws;
  CK(cudaMalloc(&dy, rows * c.cout * 2));
  CK(cudaMalloc(&x, rows * c.cin * 2));
  const size_t wn = (size_t)c.cout * c.cin * c.taps;
  CK(cudaMalloc(&dw, wn * 4));
  CK(cudaMalloc(&ws, wn * 4));
  CK(cudaMemset(ws, 0, wn * 4));
  CK(cudaMemset(dw, 0, wn * 4));
  fill_half_kernel<<<592, 256>>>(dy, rows * c.cout, 17u, 1.0f, 0);
  fill_half_kernel<<<592, 256>>>(x, rows * c.cin, 99u, 1.0f, 0);
  CK(cudaDeviceSynchronize());
  const float unscale = 0.5f;
  CKV(vnqa_conv_wgrad_h(dy, x, dw, ws, c.n_img, c.cout, c.cin, c.taps, c.flat, unscale, nullptr));
  CK(cudaDeviceSynchronize());
  const int ns = 512;
  std::vector<int> hs(3 * ns);
  for (int i = 0; i < ns; ++i) {
    hs[3 * i] = next_u32() % c.cout; hs[3 * i + 1] = next_u32() % c.cin; hs[3 * i + 2] = next_u32() % c.taps;
  }
  hs[0] = 0; hs[1] = 0; hs[2] = 0;                                   // always include the corners
  hs[3] = c.cout - 1; hs[4] = c.cin - 1; hs[5] = c.taps - 1;
  int* ds; double* dref;
  CK(cudaMalloc(&ds, hs.size() * 4));
  CK(cudaMemcpy(ds, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dref, ns * 8));
  wgrad_ref_kernel<<<ns, 256>>>(dy, x, dref, ds, ns, c.cout, c.cin, c.taps, c.n_img, c.flat);
  CK(cudaDeviceSynchronize());
  auto href = from_dev<double>(dref, ns);
  auto hw = from_dev<float>(dw, wn);
  if (getenv("VNQA_TEST_VERBOSE")) {
    size_t nz = 0; double mxv = 0;
    for (size_t i = 0; i < wn; ++i) { if (hw[i] != 0.f) ++nz; mxv = fmax(mxv, fabs(hw[i])); }
    printf("   dw: %zu of %zu non-zero, max |dw| %.4f\n", nz, wn, mxv);
  }
  double err = 0, mx = 0;
  for (int i = 0; i < ns; ++i) {
    const double got = hw[((size_t)hs[3 * i] * c.cin + hs[3 * i + 1]) * c.taps + hs[3 * i + 2]];
    const double ref = unscale * href[i];
    if (i < 4 && getenv("VNQA_TEST_VERBOSE"))
      printf("   sample co %d ci %d tap %d: got %.6f ref %.6f\n", hs[3 * i], hs[3 * i + 1], hs[3 * i + 2], got, ref);
    err = fmax(err, fabs(got - ref));
    mx = fmax(mx, fabs(ref));
  }
  const int fail = !(err / mx < (c.iters > 0 ? 2e-3 : 2e-4));   // long fp32 accumulation chains in the perf cases
  printf("%-30s taps %d cin %6d cout %5d n_img %5d  normalised max err %.3e", c.name, c.taps, c.cin, c.cout, c.n_img,
         err / mx);
  if (c.iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CKV(vnqa_conv_wgrad_h(dy, x, dw, ws, c.n_img, c.cout, c.cin, c.taps, c.flat, unscale, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.iters; ++i)
      CKV(vnqa_conv_wgrad_h(dy, x, dw, ws, c.n_img, c.cout, c.cin, c.taps, c.flat, unscale, nullptr));
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

// ---------------------------------------------------------------------------------------------- HBM kernels
static std::vector<float> host_rand(size_t n, float scale, bool relu = false) {
  std::vector<float> v(n);
  for (auto& e : v) {
    float r = ((next_u32() & 0xFFFF) / 32768.0f - 1.0f) * scale;
    e = relu ? fmaxf(r, 0.f) : r;
  }
  return v;
}
static std::vector<__half> to_half_vec(const std::vector<float>& f) {
  std::vector<__half> h(f.size());
  for (size_t i = 0; i < f.size(); ++i) h[i] = __float2half_rn(f[i]);
  return h;
}
template <class T>
static T* to_dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
static bool close_half(double got, double want, double scale) { return fabs(got - want) <= 1.1e-3 * scale + 1e-7; }

static int run_small() {
  int fails = 0;
  {  // ---- K0
    const int B = 3, C = 96, T = 5, lens[3] = {5, 3, 2};   // 96 = one full 64-channel slab + a 32-channel one
    std::vector<int> img_of(B * T, -1);
    int n_img = 0;
    for (int t = 0; t < T; ++t)
      for (int b = 0; b < B; ++b)
        if (t < lens[b]) img_of[b * T + t] = n_img++;
    auto v = host_rand((size_t)B * C * 130 * T, 1.f, true);
    float* dv = to_dev(v);
    int* dio = to_dev(img_of);
    __half* dx;
    CK(cudaMalloc(&dx, (size_t)n_img * 130 * C * 2));
    for (int simple = 0; simple < 2; ++simple) {
      if (simple) setenv("VNQA_K0_SIMPLE", "1", 1); else unsetenv("VNQA_K0_SIMPLE");
      CK(cudaMemset(dx, 0xFF, (size_t)n_img * 130 * C * 2));
      CKV(vnqa_frames_to_nhwc_h(dv, dx, dio, B, C, T, nullptr));
      CK(cudaDeviceSynchronize());
      auto hx = from_dev<__half>(dx, (size_t)n_img * 130 * C);
      int bad = 0;
      for (int b = 0; b < B; ++b)
        for (int t = 0; t < lens[b]; ++t)
          for (int c = 0; c < C; ++c)
            for (int p = 0; p < 130; ++p) {
              const float want = __half2float(__float2half_rn(v[(((size_t)b * C + c) * 130 + p) * T + t]));
              if (__half2float(hx[((size_t)img_of[b * T + t] * 130 + p) * C + c]) != want) ++bad;
            }
      printf("%-30s %s\n", simple ? "frames_to_nhwc_h (per-pixel)" : "frames_to_nhwc_h (row-tiled)", bad ? "FAIL" : "ok");
      fails += bad != 0;
    }
    unsetenv("VNQA_K0_SIMPLE");
    cudaFree(dv); cudaFree(dio); cudaFree(dx);
  }
  const int n_img = 6, C = 128, F = 3;
  const size_t n = (size_t)n_img * 130 * C;
  std::vector<int> img_frame = {0, 0, 0, 1, 1, 2}, frame_ct = {3, 2, 1};
  int* dif = to_dev(img_frame);
  int* dfc = to_dev(frame_ct);
  {  // ---- bn_apply_h
    auto y = host_rand(n, 2.f, true), sc = host_rand((size_t)F * C, 1.f), sh = host_rand((size_t)F * C, 1.f);
    float *dy = to_dev(y), *dsc = to_dev(sc), *dsh = to_dev(sh);
    __half* dx;
    CK(cudaMalloc(&dx, n * 2));
    CKV(vnqa_bn_apply_h(dy, dx, dif, dsc, dsh, n_img, C, nullptr));
    CK(cudaDeviceSynchronize());
    auto hx = from_dev<__half>(dx, n);
    int bad = 0;
    for (size_t i = 0; i < n; ++i) {
      const int c = (int)(i % C), f = img_frame[i / ((size_t)130 * C)];
      const double want = (double)y[i] * sc[(size_t)f * C + c] + sh[(size_t)f * C + c];
      if (!close_half(__half2float(hx[i]), want, fabs(want))) ++bad;
    }
    printf("%-30s %s\n", "bn_apply_h", bad ? "FAIL" : "ok");
    fails += bad != 0;
    cudaFree(dy); cudaFree(dsc); cudaFree(dsh); cudaFree(dx);
  }
  {  // ---- film_backward_h
    const int film_ld = 2 * C + 16, alpha_off = 8;
    const float unscale = 0.125f;
    auto dout = to_half_vec(host_rand(n, 1.f)), z = to_half_vec(host_rand(n, 1.f));
    auto film = host_rand((size_t)n_img * film_ld, 1.f);
    std::vector<int> film_row = {4, 2, 0, 5, 1, 3};
    __half *ddout = to_dev(dout), *dz_in = to_dev(z), *ddz;
    float *dfilm_in = to_dev(film), *ddfilm, *ddb;
    int* dfr = to_dev(film_row);
    CK(cudaMalloc(&ddz, n * 2));
    CK(cudaMalloc(&ddfilm, film.size() * 4)); CK(cudaMemset(ddfilm, 0, film.size() * 4));
    CK(cudaMalloc(&ddb, C * 4)); CK(cudaMemset(ddb, 0, C * 4));
    CKV(vnqa_film_backward_h(ddout, dz_in, dfilm_in, dfr, film_ld, alpha_off, n_img, C, ddz, ddfilm, ddb, 1, unscale, nullptr));
    CK(cudaDeviceSynchronize());
    auto hdz = from_dev<__half>(ddz, n);
    auto hdfilm = from_dev<float>(ddfilm, film.size());
    auto hdb = from_dev<float>(ddb, C);
    int bad = 0;
    std::vector<double> sdb(C, 0.0);
    for (int img = 0; img < n_img; ++img)
      for (int c = 0; c < C; ++c) {
        const size_t fr = (size_t)film_row[img] * film_ld + alpha_off;
        const double al = film[fr + c], be = film[fr + C + c];
        double da = 0, db = 0;
        for (int p = 0; p < 130; ++p) {
          const size_t i = ((size_t)img * 130 + p) * C + c;
          const double zz = __half2float(z[i]), dd = __half2float(dout[i]);
          const double d = (float)al * (float)zz + (float)be > 0.f ? dd : 0.0;
          da += d * zz; db += d;
          const double want = half_round_host(al * d);
          if (!close_half(__half2float(hdz[i]), want, fabs(want))) ++bad;
          sdb[c] += want;
        }
        const double wa = al > 0 ? da * unscale : 0.0, wb = be > 0 ? db * unscale : 0.0;
        if (fabs(hdfilm[fr + c] - wa) > 1e-4 * (1 + fabs(wa))) ++bad;
        if (fabs(hdfilm[fr + C + c] - wb) > 1e-4 * (1 + fabs(wb))) ++bad;
      }
    for (int c = 0; c < C; ++c)
      if (fabs(hdb[c] - sdb[c] * unscale) > 1e-3 * (1 + fabs(sdb[c] * unscale))) ++bad;
    printf("%-30s %s (%d mismatches)\n", "film_backward_h", bad ? "FAIL" : "ok", bad);
    fails += bad != 0;
    cudaFree(ddout); cudaFree(dz_in); cudaFree(ddz); cudaFree(dfilm_in); cudaFree(ddfilm); cudaFree(ddb); cudaFree(dfr);
  }
  {  // ---- bn_backward_h
    const float unscale = 0.25f;
    auto dx = to_half_vec(host_rand(n, 1.f));
    auto y = host_rand(n, 2.f, true);
    auto gamma = host_rand(C, 1.f);
    std::vector<float> mean((size_t)F * C), rstd((size_t)F * C);
    for (int f = 0; f < F; ++f)
      for (int c = 0; c < C; ++c) {
        double s = 0, q = 0, cnt = 0;
        for (int img = 0; img < n_img; ++img)
          if (img_frame[img] == f)
            for (int p = 0; p < 130; ++p) { const double v = y[((size_t)img * 130 + p) * C + c]; s += v; q += v * v; ++cnt; }
        const double m = s / cnt, var = q / cnt - m * m;
        mean[(size_t)f * C + c] = (float)m;
        rstd[(size_t)f * C + c] = (float)(1.0 / sqrt(var + 1e-5));
      }
    __half *ddx = to_dev(dx), *ddy;
    float *dy_in = to_dev(y), *dg = to_dev(gamma), *dm = to_dev(mean), *dr = to_dev(rstd), *ds1, *ds2, *ddb, *ddg, *ddbeta;
    CK(cudaMalloc(&ddy, n * 2));
    CK(cudaMalloc(&ds1, (size_t)F * C * 4)); CK(cudaMalloc(&ds2, (size_t)F * C * 4));
    CK(cudaMalloc(&ddb, C * 4)); CK(cudaMalloc(&ddg, C * 4)); CK(cudaMalloc(&ddbeta, C * 4));
    CK(cudaMemset(ddb, 0, C * 4)); CK(cudaMemset(ddg, 0, C * 4)); CK(cudaMemset(ddbeta, 0, C * 4));
    CKV(vnqa_bn_backward_h(ddx, dy_in, dif, dfc, F, dm, dr, dg, n_img, C, ds1, ds2, ddy, ddb, ddg, ddbeta, unscale, nullptr));
    CK(cudaDeviceSynchronize());
    auto hdy = from_dev<__half>(ddy, n);
    auto hdb = from_dev<float>(ddb, C), hdg = from_dev<float>(ddg, C), hdbeta = from_dev<float>(ddbeta, C);
    int bad = 0;
    for (int c = 0; c < C; ++c) {
      double sum_db = 0, sum_dg = 0, sum_dbeta = 0;
      for (int f = 0; f < F; ++f) {
        const double m = mean[(size_t)f * C + c], r = rstd[(size_t)f * C + c];
        double s1 = 0, s2 = 0, cnt = 0;
        for (int img = 0; img < n_img; ++img)
          if (img_frame[img] == f)
            for (int p = 0; p < 130; ++p) {
              const size_t i = ((size_t)img * 130 + p) * C + c;
              const double d = __half2float(dx[i]);
              s1 += d; s2 += d * (y[i] - m) * r; ++cnt;
            }
        sum_dg += s2; sum_dbeta += s1;
        for (int img = 0; img < n_img; ++img)
          if (img_frame[img] == f)
            for (int p = 0; p < 130; ++p) {
              const size_t i = ((size_t)img * 130 + p) * C + c;
              const double xhat = (y[i] - m) * r;
              const double want = y[i] > 0.f ? half_round_host(gamma[c] * r * (__half2float(dx[i]) - s1 / cnt - xhat * s2 / cnt)) : 0.0;
              if (!close_half(__half2float(hdy[i]), want, fabs(want) + 1e-2)) ++bad;
              sum_db += want;
            }
      }
      if (fabs(hdb[c] - sum_db * unscale) > 2e-3 * (1 + fabs(sum_db * unscale))) ++bad;
      if (fabs(hdg[c] - sum_dg * unscale) > 1e-3 * (1 + fabs(sum_dg * unscale))) ++bad;
      if (fabs(hdbeta[c] - sum_dbeta * unscale) > 1e-3 * (1 + fabs(sum_dbeta * unscale))) ++bad;
    }
    printf("%-30s %s (%d mismatches)\n", "bn_backward_h", bad ? "FAIL" : "ok", bad);
    fails += bad != 0;
    cudaFree(ddx); cudaFree(ddy); cudaFree(dy_in); cudaFree(dg); cudaFree(dm); cudaFree(dr); cudaFree(ds1); cudaFree(ds2);
    cudaFree(ddb); cudaFree(ddg); cudaFree(ddbeta);
  }
  cudaFree(dif); cudaFree(dfc);
  return fails;
}

int main(int argc, char** argv) {
  bool perf = false, do_conv = true, do_wgrad = true, do_small = true;
  const char* only = nullptr;   // --case <substring>: run only the matching cases (used under ncu)
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--perf")) perf = true;
    if (!strcmp(argv[i], "--conv-only")) { do_wgrad = false; do_small = false; }
    if (!strcmp(argv[i], "--wgrad-only")) { do_conv = false; do_small = false; }
    if (!strcmp(argv[i], "--small-only")) { do_conv = false; do_wgrad = false; }
    if (!strcmp(argv[i], "--case") && i + 1 < argc) only = argv[++i];
  }
  CKV(vnqa_device_check());
  printf("SMs: %d\n", vnqa_sm_count());
  std::vector<ConvCase> conv = {
      {"gemm1x1_plain_1tile", 1, 64, 128, 8, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"gemm1x1_plain_k256", 1, 256, 128, 8, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"conv3x3_plain", 9, 64, 128, 8, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"conv3x3_plain_ragged", 9, 128, 160, 11, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"conv1x1_bias_relu", 1, 256, 256, 19, 0, 1, VNQA_EPI_BIAS_RELU, 0, 0},
      {"conv3x3_affine", 9, 64, 96, 9, 0, 1, VNQA_EPI_BIAS_RELU_AFFINE, 0, 0},
      {"conv3x3_stats", 9, 64, 128, 21, 0, 1, VNQA_EPI_BIAS_RELU_STATS, 0, 0},
      {"conv3x3_film", 9, 128, 256, 13, 0, 1, VNQA_EPI_FILM, 0, 0},
      {"conv3x3_dgrad_res", 9, 128, 128, 10, 0, 1, VNQA_EPI_DGRAD_RES, 0, 0},
      {"flat_splitk_partial", 1, 832, 128, 300, 1, 5, VNQA_EPI_PARTIAL, 0, 0},
      {"flat_plain_tail", 1, 128, 4160, 77, 1, 1, VNQA_EPI_PLAIN, 0, 0},
      {"flat_plain_cta_pair", 1, 128, 512, 300, 1, 1, VNQA_EPI_PLAIN, 0, 0},
      {"many_tiles_persistent", 9, 64, 512, 400, 0, 1, VNQA_EPI_FILM, 20000, 0},
      {"tail_conv1x1_cout32", 1, 256, 32, 11, 0, 1, VNQA_EPI_BIAS_RELU, 0, 0},     // gp / multi-hop tails
      {"tail_dgrad_cin32", 1, 32, 256, 11, 0, 1, VNQA_EPI_PLAIN, 0, 0},            // K completed by TMA zero fill
      {"tail_conv1x1_cout8_f32out", 1, 128, 8, 11, 0, 1, VNQA_EPI_BIAS_RELU, 0, 0, 1},
      {"tail_dgrad_cin8", 1, 8, 128, 11, 0, 1, VNQA_EPI_PLAIN, 0, 0},
  };
  std::vector<WgradCase> wg = {
      {"wgrad1x1_small", 1, 128, 128, 8, 0, 0},
      {"wgrad3x3_small", 9, 128, 128, 8, 0, 0},
      {"wgrad3x3_ragged", 9, 192, 320, 11, 0, 0},
      {"wgrad3x3_narrow", 9, 64, 64, 5, 0, 0},
      {"wgrad3x3_cta_pair", 9, 256, 256, 19, 0, 0},
      {"wgrad_flat", 1, 4160, 128, 300, 1, 0},
      {"wgrad_flat_small", 1, 64, 64, 77, 1, 0},
      {"wgrad_tail_cout32", 1, 256, 32, 11, 0, 0},
  };
  if (perf) {
    conv.push_back({"P conv_init 512->1024", 9, 512, 1024, 2240, 0, 1, VNQA_EPI_BIAS_RELU_AFFINE, 4096, 5});
    conv.push_back({"P conv1x1 1024->1024", 1, 1024, 1024, 2240, 0, 1, VNQA_EPI_BIAS_RELU, 4096, 10});
    conv.push_back({"P conv3x3 film 1024", 9, 1024, 1024, 2240, 0, 1, VNQA_EPI_FILM, 4096, 5});
    conv.push_back({"P conv3x3 dgrad 1024", 9, 1024, 1024, 2240, 0, 1, VNQA_EPI_DGRAD_RES, 4096, 5});
    conv.push_back({"P fc_embed splitK", 1, 10240, 128, 2240, 1, 13, VNQA_EPI_PARTIAL, 4096, 10});
    wg.push_back({"P wgrad conv_init", 9, 512, 1024, 2240, 0, 3});
    wg.push_back({"P wgrad 1x1", 1, 1024, 1024, 2240, 0, 5});
    wg.push_back({"P wgrad 3x3", 9, 1024, 1024, 2240, 0, 3});
    wg.push_back({"P wgrad fc_embed", 1, 133120, 128, 2240, 1, 5});
  }
  int fails = 0;
  if (do_small && !only) fails += run_small();
  if (do_conv)
    for (const auto& c : conv)
      if (!only || strstr(c.name, only)) fails += run_conv(c);
  if (do_wgrad)
    for (const auto& c : wg)
      if (!only || strstr(c.name, only)) fails += run_wgrad(c);
  printf("%s (%d failing cases)\n", fails ? "FAILED" : "ALL OK", fails);
  return fails ? 1 : 0;
}
