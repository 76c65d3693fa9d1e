// Standalone GPU check + micro-benchmark of vnqa_conv_gemm (no Python): compares the tcgen05 kernel with a
// naive CUDA reference (double accumulation) on small and full-size problems and prints achieved TFLOP/s.
// Build: see Makefile target `test_conv_gemm`.  Run on the B200 box: build/test_conv_gemm [--perf]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/vnqa_b200.h"

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

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__global__ void round_kernel(float* x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = tf32_rna(x[i]);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// raw accumulator reference: acc[row][co] for conv (taps 1/9 over 10x13 images) or flat GEMM.
// mode 0: exact fp32 operands; 1: operands rounded to tf32 (rna); 2: truncated to tf32.
__global__ void ref_kernel(const float* __restrict__ w /*(cout,cin,taps) or flat (cout,K)*/, const float* __restrict__ x,
                           double* __restrict__ out, const int* __restrict__ samples, long long n_samples, int cout,
                           int cin, int taps, int n_img, int flat, int x_ld, int mode) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n_samples;
       s += (long long)gridDim.x * blockDim.x) {
    long long row;
    int co;
    if (samples) {
      row = samples[2 * s];
      co = samples[2 * s + 1];
    } else {
      row = s / cout;
      co = (int)(s % cout);
    }
    double acc = 0.0;
    if (flat) {
      for (int k = 0; k < cin; ++k) {
        float a = w[(long long)co * cin + k], b = x[row * x_ld + k];
        if (mode == 1) { a = tf32_rna(a); b = tf32_rna(b); }
        if (mode == 2) { a = tf32_trunc(a); b = tf32_trunc(b); }
        acc += (double)a * (double)b;
      }
    } else {
      const int img = (int)(row / 130), pix = (int)(row % 130), h = pix / 13, wq = pix % 13;
      for (int tap = 0; tap < taps; ++tap) {
        int dy = 0, dx = 0;
        if (taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        const int hh = h + dy, ww = wq + dx;
        if (hh < 0 || hh >= 10 || ww < 0 || ww >= 13) continue;
        const float* xp = x + ((long long)img * 130 + hh * 13 + ww) * cin;
        for (int k = 0; k < cin; ++k) {
          float a = w[((long long)co * cin + k) * taps + tap], b = xp[k];
          if (mode == 1) { a = tf32_rna(a); b = tf32_rna(b); }
          if (mode == 2) { a = tf32_trunc(a); b = tf32_trunc(b); }
          acc += (double)a * (double)b;
        }
      }
    }
    out[s] = acc;
  }
}

static double tf32_rna_host(double x) {
  float f = (float)x;
  uint32_t u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t g_seed = 12345;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f;
}
static std::vector<float> rand_vec(size_t n, float scale, bool relu = false) {
  std::vector<float> v(n);
  for (auto& e : v) {
    float r = frand() * scale;
    e = relu ? fmaxf(r, 0.f) : r;
  }
  return v;
}
template <class T>
static T* to_dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct Case {
  const char* name;
  int taps, cin, cout, n_img, flat, splits, epi;
  long long samples;  // 0 = check everything
  int iters;          // >0: time it
};

static int run_case(const Case& c) {
  const int rows_pad = (c.cout + 127) / 128 * 128;
  const long long rows = c.flat ? c.n_img : (long long)c.n_img * 130;
  const int K = c.flat ? c.cin * c.splits : c.cin;
  const int T = 4;  // frames for STATS
  std::vector<float> hw = rand_vec((size_t)c.cout * K * c.taps, 0.05f);
  std::vector<float> hx = rand_vec((size_t)rows * K, 1.0f, true);
  std::vector<float> hbias = rand_vec(c.cout, 0.5f), hscale = rand_vec(c.cout, 1.0f), hshift = rand_vec(c.cout, 1.0f);
  std::vector<float> hres = rand_vec((size_t)rows * c.cout, 1.0f), hadd = rand_vec((size_t)rows * c.cout, 1.0f);
  const int film_ld = 2 * c.cout + 16;
  std::vector<float> hfilm = rand_vec((size_t)c.n_img * film_ld, 1.0f);
  std::vector<int> hfilm_row(c.n_img), hframe(c.n_img);
  for (int i = 0; i < c.n_img; ++i) {
    hfilm_row[i] = (i * 7 + 3) % c.n_img;
    hframe[i] = (int)((long long)i * T / c.n_img);
  }
  float *dw = to_dev(hw), *dx = to_dev(hx), *dbias = to_dev(hbias), *dscale = to_dev(hscale), *dshift = to_dev(hshift);
  float *dres = to_dev(hres), *dadd = to_dev(hadd), *dfilm = to_dev(hfilm);
  int *dfilm_row = to_dev(hfilm_row), *dframe = to_dev(hframe);
  float *dwp, *dout, *dz, *ddb;
  double *dsum, *dsq;
  CK(cudaMalloc(&dwp, (size_t)c.taps * rows_pad * K * 4));
  const size_t out_elems = (size_t)rows * c.cout * (c.epi == VNQA_EPI_PARTIAL ? c.splits : 1);
  CK(cudaMalloc(&dout, out_elems * 4));
  CK(cudaMalloc(&dz, (size_t)rows * c.cout * 4));
  CK(cudaMalloc(&ddb, c.cout * 4));
  CK(cudaMalloc(&dsum, (size_t)T * c.cout * 8));
  CK(cudaMalloc(&dsq, (size_t)T * c.cout * 8));
  CK(cudaMemset(dout, 0xFF, out_elems * 4));
  CK(cudaMemset(ddb, 0, c.cout * 4));
  CK(cudaMemset(dsum, 0, (size_t)T * c.cout * 8));
  CK(cudaMemset(dsq, 0, (size_t)T * c.cout * 8));
  CKV(vnqa_pack_conv_weight(dw, dwp, c.cout, K, c.taps, rows_pad, 0, VNQA_ACT_TF32, nullptr));
  // producers round MMA operands to tf32 (rna) - emulate that for the activation input; the pack kernel rounds w
  round_kernel<<<148 * 8, 256>>>(dx, (long long)rows * K);
  CK(cudaDeviceSynchronize());

  vnqa_conv_args a;
  memset(&a, 0, sizeof(a));
  a.w_packed = dwp; a.x = dx; a.taps = c.taps; a.cin = c.cin; a.cout = c.cout; a.rows_pad = rows_pad;
  a.n_img = c.n_img; a.flat = c.flat; a.x_ld = K; a.splits = c.splits; a.epi = c.epi; a.out = dout;
  a.bias = dbias; a.scale = dscale; a.shift = dshift; a.img_frame = dframe; a.stat_sum = dsum; a.stat_sq = dsq;
  a.film = dfilm; a.film_row = dfilm_row; a.film_ld = film_ld; a.alpha_off = 8; a.res = dres; a.z_out = dz;
  a.add = dadd; a.dbias = ddb;
  CKV(vnqa_conv_gemm(&a, nullptr));
  CK(cudaDeviceSynchronize());

  // reference accumulators
  long long ns = c.samples > 0 ? c.samples : rows * c.cout;
  std::vector<int> hs;
  int* dsamp = nullptr;
  if (c.samples > 0) {
    hs.resize(2 * ns);
    for (long long i = 0; i < ns; ++i) {
      g_seed = g_seed * 1664525u + 1013904223u;
      hs[2 * i] = (int)((g_seed >> 4) % rows);
      g_seed = g_seed * 1664525u + 1013904223u;
      hs[2 * i + 1] = (int)((g_seed >> 4) % c.cout);
    }
    dsamp = to_dev(hs);
  }
  std::vector<float> hout(out_elems), hz((size_t)rows * c.cout), hdb(c.cout);
  CK(cudaMemcpy(hout.data(), dout, out_elems * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hz.data(), dz, hz.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hdb.data(), ddb, c.cout * 4, cudaMemcpyDeviceToHost));
  std::vector<double> hsum((size_t)T * c.cout), hsq((size_t)T * c.cout);
  CK(cudaMemcpy(hsum.data(), dsum, hsum.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hsq.data(), dsq, hsq.size() * 8, cudaMemcpyDeviceToHost));

  double* dref;
  CK(cudaMalloc(&dref, ns * 8));
  std::vector<double> href(ns);
  double errs[3] = {0, 0, 0}, maxref = 0;
  std::vector<double> ssum((size_t)T * c.cout, 0.0), ssq((size_t)T * c.cout, 0.0), sdb(c.cout, 0.0);
  for (int mode = 0; mode < 3; ++mode) {
    ref_kernel<<<148 * 8, 256>>>(dw, dx, dref, dsamp, ns, c.cout, K, c.taps, c.n_img, c.flat, K, mode);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(href.data(), dref, ns * 8, cudaMemcpyDeviceToHost));
    for (long long s = 0; s < ns; ++s) {
      const long long row = c.samples > 0 ? hs[2 * s] : s / c.cout;
      const int co = c.samples > 0 ? hs[2 * s + 1] : (int)(s % c.cout);
      const long long idx = row * c.cout + co;
      const int img = c.flat ? 0 : (int)(row / 130);
      double acc = href[s], expect = 0;
      switch (c.epi) {
        case VNQA_EPI_BIAS_RELU: expect = fmax(acc + hbias[co], 0.0); break;
        case VNQA_EPI_BIAS_RELU_AFFINE: expect = fmax(acc + hbias[co], 0.0) * hscale[co] + hshift[co]; break;
        case VNQA_EPI_BIAS_RELU_STATS:
          expect = fmax(acc + hbias[co], 0.0);
          if (mode == 1 && c.samples == 0) {
            ssum[(size_t)hframe[img] * c.cout + co] += expect;
            ssq[(size_t)hframe[img] * c.cout + co] += expect * expect;
          }
          break;
        case VNQA_EPI_FILM: {
          const float* fr = &hfilm[(size_t)hfilm_row[img] * film_ld + 8];
          const double z = acc + hbias[co];
          expect = fmax(fr[co] * z + fr[c.cout + co], 0.0) + hres[idx];
          if (mode == 1) {
            const double ez = fabs(hz[idx] - z);
            if (ez > 1e-4 * (1 + fabs(z))) { printf("  z_out mismatch at row %lld co %d: %f vs %f\n", row, co, hz[idx], z); return 1; }
          }
          break;
        }
        case VNQA_EPI_PLAIN: expect = acc; break;
        case VNQA_EPI_DGRAD_RES:
          expect = hres[idx] > 0.f ? acc + hadd[idx] : 0.0;
          if (mode == 1 && c.samples == 0) sdb[co] += expect;
          break;
        case VNQA_EPI_PARTIAL: expect = acc; break;
      }
      // epilogues that feed another MMA store tf32-rounded values
      if (c.epi == VNQA_EPI_BIAS_RELU || c.epi == VNQA_EPI_BIAS_RELU_AFFINE || c.epi == VNQA_EPI_FILM ||
          c.epi == VNQA_EPI_DGRAD_RES)
        expect = tf32_rna_host(expect);
      double got;
      if (c.epi == VNQA_EPI_PARTIAL) {
        got = 0;
        for (int sp = 0; sp < c.splits; ++sp) got += hout[(size_t)sp * rows * c.cout + idx];
      } else {
        got = hout[idx];
      }
      const double e = fabs(got - expect);
      if (e > errs[mode]) errs[mode] = e;
      if (mode == 0 && fabs(expect) > maxref) maxref = fabs(expect);
    }
  }
  int fail = 0;
  const double n0 = errs[0] / maxref, n1 = errs[1] / maxref, n2 = errs[2] / maxref;
  printf("%-28s taps %d cin %5d cout %5d n_img %5d  normalised max err: vs fp32 %.3e  vs tf32-rna %.3e  vs tf32-trunc %.3e",
         c.name, c.taps, K, c.cout, c.n_img, n0, n1, n2);
  const bool rounded = c.epi == VNQA_EPI_BIAS_RELU || c.epi == VNQA_EPI_BIAS_RELU_AFFINE || c.epi == VNQA_EPI_FILM ||
                       c.epi == VNQA_EPI_DGRAD_RES;
  // operands are pre-rounded (rna), so the rna reference must match to accumulation-order noise; rounded outputs
  // may differ by one tf32 ulp (2^-10 relative) where the fp32 sums differ in the last bits
  if (!(n0 < 3e-3) || !(n1 < (rounded ? 1.1e-3 : 1e-4))) fail = 1;
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
      es = fmax(es, fabs(sdb[i] - hdb[i]));
      ms = fmax(ms, fabs(sdb[i]));
    }
    printf("  dbias err %.3e", es / ms);
    if (!(es / ms < 2e-3)) fail = 1;
  }
  if (c.iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) CKV(vnqa_conv_gemm(&a, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.iters; ++i) CKV(vnqa_conv_gemm(&a, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= c.iters;
    const double flop = 2.0 * rows * c.cout * (double)K * c.taps;
    printf("  | %.3f ms  %.1f TFLOP/s", ms, flop / ms * 1e-9);
  }
  printf("  %s\n", fail ? "FAIL" : "ok");
  fflush(stdout);
  cudaFree(dw); cudaFree(dx); cudaFree(dbias); cudaFree(dscale); cudaFree(dshift); cudaFree(dres); cudaFree(dadd);
  cudaFree(dfilm); cudaFree(dfilm_row); cudaFree(dframe); cudaFree(dwp); cudaFree(dout); cudaFree(dz); cudaFree(ddb);
  cudaFree(dsum); cudaFree(dsq); cudaFree(dref);
  if (dsamp) cudaFree(dsamp);
  return fail;
}

int main(int argc, char** argv) {
  bool perf = false, extended = false;
  const char* only = nullptr;   // --case <substring>: run only the matching cases (used under ncu)
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--perf")) perf = true;
    if (!strcmp(argv[i], "--extended")) extended = true;
    if (!strcmp(argv[i], "--case") && i + 1 < argc) only = argv[++i];
  }
  CKV(vnqa_device_check());
  printf("SMs: %d\n", vnqa_sm_count());
  std::vector<Case> cases = {
      {"gemm1x1_plain_1tile", 1, 32, 128, 8, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"gemm1x1_plain_k128", 1, 128, 128, 8, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"conv3x3_plain", 9, 64, 128, 8, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"conv3x3_plain_ragged", 9, 96, 160, 11, 0, 1, VNQA_EPI_PLAIN, 0, 0},
      {"conv1x1_bias_relu", 1, 256, 256, 19, 0, 1, VNQA_EPI_BIAS_RELU, 0, 0},
      {"conv3x3_affine", 9, 64, 96, 9, 0, 1, VNQA_EPI_BIAS_RELU_AFFINE, 0, 0},
      {"conv3x3_stats", 9, 64, 128, 21, 0, 1, VNQA_EPI_BIAS_RELU_STATS, 0, 0},
      {"conv3x3_film", 9, 128, 256, 13, 0, 1, VNQA_EPI_FILM, 0, 0},
      {"conv3x3_dgrad_res", 9, 128, 128, 10, 0, 1, VNQA_EPI_DGRAD_RES, 0, 0},
      {"flat_splitk_partial", 1, 832, 128, 300, 1, 5, VNQA_EPI_PARTIAL, 0, 0},
      {"flat_plain_tail", 1, 128, 4160, 77, 1, 1, VNQA_EPI_PLAIN, 0, 0},
      {"many_tiles_persistent", 9, 64, 512, 400, 0, 1, VNQA_EPI_FILM, 20000, 0},
  };
  if (extended) {   // added after the last GPU session of round 1: not yet run on hardware
    cases.push_back({"flat_plain_cta_pair", 1, 128, 512, 300, 1, 1, VNQA_EPI_PLAIN, 0, 0});
    cases.push_back({"conv3x3_film_cta_pair_ragged", 9, 64, 256, 21, 0, 1, VNQA_EPI_FILM, 0, 0});
  }
  if (perf) {
    cases.push_back({"P conv_init 512->1024", 9, 512, 1024, 2240, 0, 1, VNQA_EPI_BIAS_RELU_AFFINE, 4096, 5});
    cases.push_back({"P conv1x1 1024->1024", 1, 1024, 1024, 2240, 0, 1, VNQA_EPI_BIAS_RELU, 4096, 10});
    cases.push_back({"P conv3x3 film 1024", 9, 1024, 1024, 2240, 0, 1, VNQA_EPI_FILM, 4096, 5});
    cases.push_back({"P fc_embed splitK", 1, 10240, 128, 2240, 1, 13, VNQA_EPI_PARTIAL, 4096, 10});
  }
  int fails = 0;
  for (const auto& c : cases)
    if (!only || strstr(c.name, only)) fails += run_case(c);
  printf("%s (%d failing cases)\n", fails ? "FAILED" : "ALL OK", fails);
  return fails ? 1 : 0;
}
