// K7: temporal attention read-out of FiLMAttnPretrainedStem (models/film_attn_pt_stem.py:253-295) and its
// backward.  One CTA per clip (4*A threads): frame scores, mask, then T iterations of
// { softmax over frames, weighted pooling, LSTMCell }.  Everything a clip needs (T x A features = 17.9 KB) sits
// in shared memory; reductions over frames use warp shuffles.  fc_hidden_attn(h) is constant along T and cancels
// in the softmax (quirk Q4) - it is still evaluated so that results and ~0 gradients match the reference.
#include "common.h"

namespace vnqa {

constexpr float NEG_MASK = -2147483648.f;  // -(1 << 31), models/film_attn_pt_stem.py:251
constexpr int MAX_T = 64;

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct TailParams {
  const float* feats;   // [B][T][A], zero where the clip has no frame
  const int* v_lens;    // [B]
  int n_proc;           // frames processed = v_lens[0]
  int T, A;
  const float *w1, *b1, *wh, *bh;           // fc_attn_1, fc_hidden_attn
  const float *wih_t, *whh_t, *b_ih, *b_hh; // LSTMCell, weights transposed [A][4A]
  float* hs;            // [B][T][A] outputs
  float *coefs, *ctxt, *gates, *cs;  // saved for backward (may be null): [B][T][T], [B][T][A], [B][T][4A], [B][T][A]
};

__global__ void attn_tail_fwd_kernel(const TailParams p) {
  extern __shared__ float sm[];
  const int T = p.T, A = p.A, A4 = 4 * A;
  float* F = sm;               // [T][A]
  float* s = F + T * A;        // [MAX_T] scores
  float* m = s + MAX_T;        // [MAX_T] masks
  float* coef = m + MAX_T;     // [MAX_T]
  float* h = coef + MAX_T;     // [A]
  float* cx = h + A;           // [A] context
  float* g = cx + A;           // [4A]
  float* red = g + A4;         // [2]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int vlen = p.v_lens[b];
  for (int i = tid; i < T * A; i += blockDim.x) F[i] = p.feats[(long long)b * T * A + i];
  if (tid < A) h[tid] = 0.f;
  float c = 0.f;
  __syncthreads();
  // frame scores on valid (b,t) only, 0 elsewhere (:268-281); mask only for processed frames (:251, Q5)
  for (int t = warp; t < T; t += nwarp) {
    float acc = 0.f;
    for (int a = lane; a < A; a += 32) acc = fmaf(F[t * A + a], p.w1[a], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      s[t] = t < vlen ? acc + p.b1[0] : 0.f;
      m[t] = (t >= vlen && t < p.n_proc) ? NEG_MASK : 0.f;
    }
  }
  __syncthreads();
  for (int it = 0; it < T; ++it) {
    if (warp == 0) {
      float acc = 0.f;
      for (int a = lane; a < A; a += 32) acc = fmaf(h[a], p.wh[a], acc);
      const float v = warp_sum(acc) + p.bh[0];
      // softmax over frames of (v + s) + m, two frames per lane (T <= 64)
      const float l0 = lane < T ? (v + s[lane]) + m[lane] : -INFINITY;
      const float l1 = lane + 32 < T ? (v + s[lane + 32]) + m[lane + 32] : -INFINITY;
      const float mx = warp_max(fmaxf(l0, l1));
      const float e0 = lane < T ? expf(l0 - mx) : 0.f, e1 = lane + 32 < T ? expf(l1 - mx) : 0.f;
      const float inv = 1.f / warp_sum(e0 + e1);
      if (lane < T) coef[lane] = e0 * inv;
      if (lane + 32 < T) coef[lane + 32] = e1 * inv;
    }
    __syncthreads();
    if (tid < A) {
      float acc = 0.f;
      for (int t = 0; t < T; ++t) acc = fmaf(coef[t], F[t * A + tid], acc);
      cx[tid] = acc;
    }
    __syncthreads();
    {
      float a1 = p.b_ih[tid], a2 = p.b_hh[tid];
      for (int k = 0; k < A; ++k) {
        a1 = fmaf(__ldg(p.wih_t + (long long)k * A4 + tid), cx[k], a1);
        a2 = fmaf(__ldg(p.whh_t + (long long)k * A4 + tid), h[k], a2);
      }
      g[tid] = a1 + a2;
    }
    __syncthreads();
    const long long row = (long long)b * T + it;
    if (tid < A) {
      const float ig = sigm(g[tid]), fg = sigm(g[A + tid]), gg = tanhf(g[2 * A + tid]), og = sigm(g[3 * A + tid]);
      c = fg * c + ig * gg;
      const float hn = og * tanhf(c);
      h[tid] = hn;
      p.hs[row * A + tid] = hn;
      if (p.gates) {
        p.gates[row * A4 + tid] = ig;
        p.gates[row * A4 + A + tid] = fg;
        p.gates[row * A4 + 2 * A + tid] = gg;
        p.gates[row * A4 + 3 * A + tid] = og;
        p.cs[row * A + tid] = c;
        p.ctxt[row * A + tid] = cx[tid];
      }
    }
    if (p.coefs && tid < T) p.coefs[row * T + tid] = coef[tid];
    __syncthreads();
  }
  (void)red;
}

struct TailBwdParams {
  const float* feats; const int* v_lens; int n_proc, T, A;
  const float *w1, *wh, *wih, *whh;   // wih/whh in PyTorch layout [4A][A]
  const float *hs, *coefs, *gates, *cs;
  const float* dhs;        // [B][T][A]
  float* dfeats;           // [B][T][A]
  float* dgates;           // [B][T][4A] pre-activation gate gradients (dW_ih/dW_hh/db by GEMM afterwards)
  float *dw1, *db1, *dwh, *dbh;  // atomically accumulated
};

__global__ void attn_tail_bwd_kernel(const TailBwdParams p) {
  extern __shared__ float sm[];
  const int T = p.T, A = p.A, A4 = 4 * A;
  float* F = sm;                 // [T][A]
  float* dF = F + T * A;         // [T][A]
  float* ds = dF + T * A;        // [MAX_T]
  float* dcoef = ds + MAX_T;     // [MAX_T]
  float* coef = dcoef + MAX_T;   // [MAX_T]
  float* dG = coef + MAX_T;      // [4A]
  float* part = dG + A4;         // [2][4A] partial sums for dctxt / dh_prev
  float* dcx = part + 2 * A4;    // [A]
  float* dhn = dcx + A;          // [A] dh flowing to the previous iteration
  float* hprev = dhn + A;        // [A]
  float* sc = hprev + A;         // [4] scalars
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int qd = tid / A, k = tid % A;
  const int vlen = p.v_lens[b];
  for (int i = tid; i < T * A; i += blockDim.x) {
    F[i] = p.feats[(long long)b * T * A + i];
    dF[i] = 0.f;
  }
  if (tid < MAX_T) ds[tid] = 0.f;
  if (tid < A) dhn[tid] = 0.f;
  float dc = 0.f, dwh_acc = 0.f, dbh_acc = 0.f;
  __syncthreads();
  for (int it = T - 1; it >= 0; --it) {
    const long long row = (long long)b * T + it;
    if (tid < A) {
      const float dh = p.dhs[row * A + tid] + dhn[tid];
      const float ig = p.gates[row * A4 + tid], fg = p.gates[row * A4 + A + tid], gg = p.gates[row * A4 + 2 * A + tid],
                  og = p.gates[row * A4 + 3 * A + tid];
      const float cc = p.cs[row * A + tid];
      const float cprev = it > 0 ? p.cs[(row - 1) * A + tid] : 0.f;
      const float tc = tanhf(cc);
      const float dct = dc + dh * og * (1.f - tc * tc);
      dG[tid] = dct * gg * ig * (1.f - ig);
      dG[A + tid] = dct * cprev * fg * (1.f - fg);
      dG[2 * A + tid] = dct * ig * (1.f - gg * gg);
      dG[3 * A + tid] = dh * tc * og * (1.f - og);
      dc = dct * fg;
      hprev[tid] = it > 0 ? p.hs[(row - 1) * A + tid] : 0.f;
    }
    if (tid < T) coef[tid] = p.coefs[row * T + tid];
    __syncthreads();
    p.dgates[row * A4 + tid] = dG[tid];
    {
      float a1 = 0.f, a2 = 0.f;
      for (int jj = 0; jj < A; ++jj) {
        const float d = dG[qd * A + jj];
        a1 = fmaf(__ldg(p.wih + (long long)(qd * A + jj) * A + k), d, a1);
        a2 = fmaf(__ldg(p.whh + (long long)(qd * A + jj) * A + k), d, a2);
      }
      part[tid] = a1;
      part[A4 + tid] = a2;
    }
    __syncthreads();
    if (tid < A) {
      dcx[tid] = part[tid] + part[A + tid] + part[2 * A + tid] + part[3 * A + tid];
      dhn[tid] = part[A4 + tid] + part[A4 + A + tid] + part[A4 + 2 * A + tid] + part[A4 + 3 * A + tid];
    }
    __syncthreads();
    for (int t = warp; t < T; t += nwarp) {
      float acc = 0.f;
      for (int a = lane; a < A; a += 32) acc = fmaf(F[t * A + a], dcx[a], acc);
      acc = warp_sum(acc);
      if (lane == 0) dcoef[t] = acc;
    }
    for (int i = tid; i < T * A; i += blockDim.x) dF[i] = fmaf(coef[i / A], dcx[i % A], dF[i]);
    __syncthreads();
    if (warp == 0) {
      const float c0 = lane < T ? coef[lane] : 0.f, c1 = lane + 32 < T ? coef[lane + 32] : 0.f;
      const float d0 = lane < T ? dcoef[lane] : 0.f, d1 = lane + 32 < T ? dcoef[lane + 32] : 0.f;
      const float dot = warp_sum(c0 * d0 + c1 * d1);
      const float l0 = c0 * (d0 - dot), l1 = c1 * (d1 - dot);
      if (lane < T) ds[lane] += l0;
      if (lane + 32 < T) ds[lane + 32] += l1;
      const float dv = warp_sum(l0 + l1);  // ~0: the hidden-state logit is constant along T (Q4)
      if (lane == 0) sc[0] = dv;
    }
    __syncthreads();
    if (tid < A) {
      const float dv = sc[0];
      dwh_acc = fmaf(dv, hprev[tid], dwh_acc);
      dhn[tid] = fmaf(dv, p.wh[tid], dhn[tid]);
    }
    if (tid == 0) dbh_acc += sc[0];
    __syncthreads();
  }
  // frame-score layer: s[t] = w1 . F[t] + b1 on valid frames
  float dw1_acc = 0.f;
  if (tid < A) {
    for (int t = 0; t < vlen && t < T; ++t) {
      dF[t * A + tid] = fmaf(ds[t], p.w1[tid], dF[t * A + tid]);
      dw1_acc = fmaf(ds[t], F[t * A + tid], dw1_acc);
    }
    atomicAdd(p.dw1 + tid, dw1_acc);
    atomicAdd(p.dwh + tid, dwh_acc);
  }
  if (tid == 0) {
    float d = 0.f;
    for (int t = 0; t < vlen && t < T; ++t) d += ds[t];
    atomicAdd(p.db1, d);
    atomicAdd(p.dbh, dbh_acc);
  }
  __syncthreads();
  for (int i = tid; i < T * A; i += blockDim.x) p.dfeats[(long long)b * T * A + i] = dF[i];
}

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_attn_tail_forward(const float* feats, const int32_t* v_lens, int32_t n_proc, int32_t B, int32_t T,
                                      int32_t A, const float* w1, const float* b1, const float* wh, const float* bh,
                                      const float* wih_t, const float* whh_t, const float* b_ih, const float* b_hh,
                                      float* hs, float* coefs, float* ctxt, float* gates, float* cs, void* stream) {
  VNQA_REQUIRE(feats && v_lens && w1 && b1 && wh && bh && wih_t && whh_t && b_ih && b_hh && hs, "vnqa_attn_tail_forward: null");
  VNQA_REQUIRE(T > 0 && T <= MAX_T && A % 32 == 0 && A > 0 && A <= 256, "vnqa_attn_tail_forward: bad T/A (%d, %d)", T, A);
  TailParams p{feats, v_lens, n_proc, T, A, w1, b1, wh, bh, wih_t, whh_t, b_ih, b_hh, hs, coefs, ctxt, gates, cs};
  const int smem = (T * A + 3 * MAX_T + 2 * A + 4 * A + 2) * 4;
  VNQA_CHECK_CUDA(cudaFuncSetAttribute(attn_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attn_tail_fwd_kernel<<<B, 4 * A, smem, as_stream(stream)>>>(p);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_attn_tail_backward(const float* feats, const int32_t* v_lens, int32_t n_proc, int32_t B, int32_t T,
                                       int32_t A, const float* w1, const float* wh, const float* wih, const float* whh,
                                       const float* hs, const float* coefs, const float* gates, const float* cs,
                                       const float* dhs, float* dfeats, float* dgates, float* dw1, float* db1,
                                       float* dwh, float* dbh, void* stream) {
  VNQA_REQUIRE(feats && v_lens && w1 && wh && wih && whh && hs && coefs && gates && cs && dhs && dfeats && dgates &&
                   dw1 && db1 && dwh && dbh, "vnqa_attn_tail_backward: null");
  VNQA_REQUIRE(T > 0 && T <= MAX_T && A % 32 == 0 && A > 0 && A <= 256, "vnqa_attn_tail_backward: bad T/A");
  TailBwdParams p{feats, v_lens, n_proc, T, A, w1, wh, wih, whh, hs, coefs, gates, cs, dhs, dfeats, dgates, dw1, db1, dwh, dbh};
  const int smem = (2 * T * A + 3 * MAX_T + 4 * A + 8 * A + 3 * A + 4) * 4;
  VNQA_CHECK_CUDA(cudaFuncSetAttribute(attn_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attn_tail_bwd_kernel<<<B, 4 * A, smem, as_stream(stream)>>>(p);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
