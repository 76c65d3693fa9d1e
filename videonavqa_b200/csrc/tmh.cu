// K9: TimeMultiHopFiLMPretrainedStem FiLM decoder (models/time_multi_hop_pt_stem.py:124-184) and its backward.
//
// Per frame image (frame i, clip b < ct_i) the reference computes, block after block ("hops"):
//   e_w = h (.) s_w ;  a = softmax_w(w_att . e_w + b_att) over ALL W = max(q_len) word slots - padded slots are NOT
//   masked (quirk Q10: their states are zero, so their logit is the bias) ;  h' = sum_w a_w e_w ;
//   film_k = LayerNorm_{2*C*NB}(W_out h' + b_out)[2kC : 2(k+1)C]
// with h initialised to LayerNorm_H(last question-LSTM state) and chained across blocks.  The hop chain is tiny
// (W <= 56 words x H = 128) and latency-bound: one CTA per frame image keeps the word states in shared memory and
// runs all hops; the Linear is one small SGEMM per hop (vnqa_sgemm); the full-width LayerNorm is a block-per-row
// kernel that only writes the 2C-wide slice a block uses.
#include "common.h"

namespace vnqa {

constexpr int TMH_MAX_W = 64;
constexpr float LN_EPS = 1e-5f;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum (all threads get the result); red must hold >= 32 floats
__device__ __forceinline__ float bsum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = wsum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// ------------------------------------------------------------------------------------------------ hop chain
struct HopFwd {
  const float* states;   // [n_pass][B][L][H] question-LSTM outputs (zero beyond q_len)
  const float* h_last;   // [n_pass][B][H]
  const int* img_frame;  // [n_img]
  const int* img_clip;   // [n_img]
  int n_img, B, L, H, W, NB;
  const float *enc_g, *enc_b;   // encoder_norm
  const float *w_att, *b_att;   // fc_hidden_attn (H -> 1)
  float* hvec;           // [NB][n_img][H]  attended vectors = inputs of fc_attn_out
  float* enc_out;        // [n_img][H]      LayerNorm(h_last) = context of hop 0
  float* enc_stat;       // [n_img][2]      mean, rstd
  float* coefs;          // [NB][n_img][W]  attention weights
};

// blockDim.x = H (multiple of 32, <= 256)
__global__ void tmh_hops_fwd_kernel(const HopFwd p) {
  extern __shared__ float sm[];
  const int H = p.H, W = p.W;
  float* S = sm;                 // [W][H]
  float* hc = S + W * H;         // [H] current context
  float* lg = hc + H;            // [TMH_MAX_W]
  float* red = lg + TMH_MAX_W;   // [32]
  const int img = blockIdx.x, j = threadIdx.x, lane = j & 31, warp = j >> 5, nw = blockDim.x >> 5;
  const int fr = p.img_frame[img], b = p.img_clip[img];
  const float* sp = p.states + ((long long)fr * p.B + b) * p.L * H;
  for (int w = 0; w < W; ++w) S[w * H + j] = sp[(long long)w * H + j];
  const float x = p.h_last[((long long)fr * p.B + b) * H + j];
  const float mean = bsum(x, red) / H;
  const float d = x - mean;
  const float var = bsum(d * d, red) / H;
  const float rstd = rsqrtf(var + LN_EPS);
  const float enc = d * rstd * p.enc_g[j] + p.enc_b[j];
  p.enc_out[(long long)img * H + j] = enc;
  if (j == 0) { p.enc_stat[img * 2] = mean; p.enc_stat[img * 2 + 1] = rstd; }
  hc[j] = enc;
  const float wa = p.w_att[j], ba = p.b_att[0];
  for (int k = 0; k < p.NB; ++k) {
    __syncthreads();
    for (int w = warp; w < W; w += nw) {     // logit_w = sum_c w_att[c] * h[c] * s_w[c] + b
      float acc = 0.f;
      for (int c = lane; c < H; c += 32) acc = fmaf(p.w_att[c] * hc[c], S[w * H + c], acc);
      acc = wsum(acc);
      if (lane == 0) lg[w] = acc + ba;
    }
    __syncthreads();
    if (warp == 0) {                          // softmax over the W word slots (W <= 64)
      const float l0 = lane < W ? lg[lane] : -INFINITY, l1 = lane + 32 < W ? lg[lane + 32] : -INFINITY;
      const float mx = wmax(fmaxf(l0, l1));
      const float e0 = lane < W ? expf(l0 - mx) : 0.f, e1 = lane + 32 < W ? expf(l1 - mx) : 0.f;
      const float inv = 1.f / wsum(e0 + e1);
      if (lane < W) lg[lane] = e0 * inv;
      if (lane + 32 < W) lg[lane + 32] = e1 * inv;
    }
    __syncthreads();
    float acc = 0.f;
    for (int w = 0; w < W; ++w) acc = fmaf(lg[w], S[w * H + j], acc);
    const float hv = hc[j] * acc;             // h'[j] = sum_w a_w * h[j] * s_w[j]
    p.hvec[((long long)k * p.n_img + img) * H + j] = hv;
    if (j < W) p.coefs[((long long)k * p.n_img + img) * W + j] = lg[j];
    if (j + blockDim.x < W) p.coefs[((long long)k * p.n_img + img) * W + j + blockDim.x] = lg[j + blockDim.x];
    __syncthreads();
    hc[j] = hv;
  }
  (void)wa;
}

struct HopBwd {
  const float* states; const float* h_last; const int* img_frame; const int* img_clip; const int* q_lens;
  int n_img, B, L, H, W, NB;
  const float *enc_g, *w_att;
  const float *hvec, *enc_out, *enc_stat, *coefs;
  const float* dhvec;    // [NB][n_img][H] gradient of the attended vectors (from fc_attn_out)
  float* dstates;        // [n_pass][B][L][H] gradient of the LSTM outputs (each image owns its slice)
  float *dw_att, *db_att, *denc_g, *denc_b;   // atomically accumulated
};

__global__ void tmh_hops_bwd_kernel(const HopBwd p) {
  extern __shared__ float sm[];
  const int H = p.H, W = p.W;
  float* S = sm;                 // [W][H]
  float* dS = S + W * H;         // [W][H]
  float* hcs = dS + W * H;       // [H]
  float* dhs = hcs + H;          // [H]
  float* cf = dhs + H;           // [TMH_MAX_W] coefs
  float* dl = cf + TMH_MAX_W;    // [TMH_MAX_W] dcoef -> dlogit
  float* red = dl + TMH_MAX_W;   // [32]
  const int img = blockIdx.x, j = threadIdx.x, lane = j & 31, warp = j >> 5, nw = blockDim.x >> 5;
  const int fr = p.img_frame[img], b = p.img_clip[img];
  const long long sbase = ((long long)fr * p.B + b) * p.L * H;
  for (int w = 0; w < W; ++w) { S[w * H + j] = p.states[sbase + (long long)w * H + j]; dS[w * H + j] = 0.f; }
  const float wa = p.w_att[j];
  float carry = 0.f, dwa = 0.f, dba = 0.f;
  for (int k = p.NB - 1; k >= 0; --k) {
    const float hc = k == 0 ? p.enc_out[(long long)img * H + j] : p.hvec[((long long)(k - 1) * p.n_img + img) * H + j];
    const float dh = p.dhvec[((long long)k * p.n_img + img) * H + j] + carry;
    __syncthreads();
    hcs[j] = hc; dhs[j] = dh;
    if (j < W) cf[j] = p.coefs[((long long)k * p.n_img + img) * W + j];
    if (j + blockDim.x < W) cf[j + blockDim.x] = p.coefs[((long long)k * p.n_img + img) * W + j + blockDim.x];
    __syncthreads();
    for (int w = warp; w < W; w += nw) {      // dcoef_w = sum_c e_w[c] * dh[c]
      float acc = 0.f;
      for (int c = lane; c < H; c += 32) acc = fmaf(hcs[c] * S[w * H + c], dhs[c], acc);
      acc = wsum(acc);
      if (lane == 0) dl[w] = acc;
    }
    __syncthreads();
    if (warp == 0) {                           // softmax backward
      const float c0 = lane < W ? cf[lane] : 0.f, c1 = lane + 32 < W ? cf[lane + 32] : 0.f;
      const float d0 = lane < W ? dl[lane] : 0.f, d1 = lane + 32 < W ? dl[lane + 32] : 0.f;
      const float dot = wsum(c0 * d0 + c1 * d1);
      const float g0 = c0 * (d0 - dot), g1 = c1 * (d1 - dot);
      if (lane < W) dl[lane] = g0;
      if (lane + 32 < W) dl[lane + 32] = g1;
      const float s = wsum(g0 + g1);
      if (lane == 0) red[31] = s;
    }
    __syncthreads();
    if (j == 0) dba += red[31];
    float dhc = 0.f;
    for (int w = 0; w < W; ++w) {
      const float s = S[w * H + j];
      const float de = fmaf(cf[w], dh, dl[w] * wa);   // d e_w[j]
      dhc = fmaf(de, s, dhc);
      dS[w * H + j] = fmaf(de, hc, dS[w * H + j]);
      dwa = fmaf(dl[w] * s, hc, dwa);
    }
    carry = dhc;
  }
  // encoder LayerNorm backward; its input is the LSTM state at the clip's last token
  const float x = p.h_last[((long long)fr * p.B + b) * H + j];
  const float mean = p.enc_stat[img * 2], rstd = p.enc_stat[img * 2 + 1];
  const float xhat = (x - mean) * rstd;
  const float dxh = carry * p.enc_g[j];
  const float m1 = bsum(dxh, red) / H;
  const float m2 = bsum(dxh * xhat, red) / H;
  const float dx = rstd * (dxh - m1 - xhat * m2);
  atomicAdd(p.denc_g + j, carry * xhat);
  atomicAdd(p.denc_b + j, carry);
  atomicAdd(p.dw_att + j, dwa);
  if (j == 0) atomicAdd(p.db_att, dba);
  const int last = p.q_lens[b] - 1;
  for (int w = 0; w < W; ++w)
    p.dstates[sbase + (long long)w * H + j] = dS[w * H + j] + (w == last ? dx : 0.f);
}

// ------------------------------------------------------------------------------------------------ LayerNorm slice
// cond [rows][N] -> stats [rows][2]; film[film_row[r]*film_ld + off + c] = LN(cond[r])[off + c] for c in [0, width)
__global__ void __launch_bounds__(256) ln_slice_fwd_kernel(const float* __restrict__ cond, int N,
                                                           const float* __restrict__ g, const float* __restrict__ bta,
                                                           const int* __restrict__ film_row, int film_ld, int off,
                                                           int width, float* __restrict__ film,
                                                           float* __restrict__ stat) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float* x = cond + (long long)r * N;
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += x[i];
  const float mean = bsum(s, red) / N;
  float v = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { const float d = x[i] - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(bsum(v, red) / N + LN_EPS);
  if (threadIdx.x == 0) { stat[r * 2] = mean; stat[r * 2 + 1] = rstd; }
  float* o = film + (long long)film_row[r] * film_ld + off;
  for (int c = threadIdx.x; c < width; c += blockDim.x)
    o[c] = (x[off + c] - mean) * rstd * g[off + c] + bta[off + c];
}

// upstream gradient lives only on the slice [off, off+width) of every row (dfilm rows); dcond is dense.
__global__ void __launch_bounds__(256) ln_slice_bwd_kernel(const float* __restrict__ cond, int N,
                                                           const float* __restrict__ g,
                                                           const int* __restrict__ film_row, int film_ld, int off,
                                                           int width, const float* __restrict__ dfilm,
                                                           const float* __restrict__ stat, float* __restrict__ dcond,
                                                           float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float* x = cond + (long long)r * N;
  const float mean = stat[r * 2], rstd = stat[r * 2 + 1];
  const float* up = dfilm + (long long)film_row[r] * film_ld + off;
  float a = 0.f, b2 = 0.f;
  for (int c = threadIdx.x; c < width; c += blockDim.x) {
    const float gu = up[c];
    const float xh = (x[off + c] - mean) * rstd;
    const float t = gu * g[off + c];
    a += t;
    b2 = fmaf(t, xh, b2);
    atomicAdd(dg + off + c, gu * xh);
    atomicAdd(db + off + c, gu);
  }
  const float m1 = bsum(a, red) / N, m2 = bsum(b2, red) / N;
  float* o = dcond + (long long)r * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float xh = (x[i] - mean) * rstd;
    const float t = (i >= off && i < off + width) ? up[i - off] * g[i] : 0.f;
    o[i] = rstd * (t - m1 - xh * m2);
  }
}

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_tmh_hops_forward(const float* states, const float* h_last, const int32_t* img_frame,
                                     const int32_t* img_clip, int32_t n_img, int32_t B, int32_t L, int32_t H, int32_t W,
                                     int32_t NB, const float* enc_g, const float* enc_b, const float* w_att,
                                     const float* b_att, float* hvec, float* enc_out, float* enc_stat, float* coefs,
                                     void* stream) {
  VNQA_REQUIRE(states && h_last && img_frame && img_clip && hvec && enc_out && enc_stat && coefs, "vnqa_tmh_hops_forward: null");
  VNQA_REQUIRE(H % 32 == 0 && H <= 256 && W >= 1 && W <= TMH_MAX_W && W <= L, "vnqa_tmh_hops_forward: bad H/W (%d, %d)", H, W);
  HopFwd p{states, h_last, img_frame, img_clip, n_img, B, L, H, W, NB, enc_g, enc_b, w_att, b_att, hvec, enc_out, enc_stat, coefs};
  const int smem = (W * H + H + TMH_MAX_W + 32) * 4;
  VNQA_CHECK_CUDA(cudaFuncSetAttribute(tmh_hops_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tmh_hops_fwd_kernel<<<n_img, H, smem, as_stream(stream)>>>(p);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_tmh_hops_backward(const float* states, const float* h_last, const int32_t* img_frame,
                                      const int32_t* img_clip, const int32_t* q_lens, int32_t n_img, int32_t B, int32_t L,
                                      int32_t H, int32_t W, int32_t NB, const float* enc_g, const float* w_att,
                                      const float* hvec, const float* enc_out, const float* enc_stat, const float* coefs,
                                      const float* dhvec, float* dstates, float* dw_att, float* db_att, float* denc_g,
                                      float* denc_b, void* stream) {
  VNQA_REQUIRE(states && h_last && img_frame && img_clip && q_lens && hvec && enc_out && enc_stat && coefs && dhvec &&
                   dstates && dw_att && db_att && denc_g && denc_b, "vnqa_tmh_hops_backward: null");
  VNQA_REQUIRE(H % 32 == 0 && H <= 256 && W >= 1 && W <= TMH_MAX_W && W <= L, "vnqa_tmh_hops_backward: bad H/W");
  HopBwd p{states, h_last, img_frame, img_clip, q_lens, n_img, B, L, H, W, NB, enc_g, w_att, hvec, enc_out, enc_stat,
           coefs, dhvec, dstates, dw_att, db_att, denc_g, denc_b};
  const int smem = (2 * W * H + 2 * H + 2 * TMH_MAX_W + 32) * 4;
  VNQA_CHECK_CUDA(cudaFuncSetAttribute(tmh_hops_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tmh_hops_bwd_kernel<<<n_img, H, smem, as_stream(stream)>>>(p);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_ln_slice_forward(const float* cond, int32_t rows, int32_t N, const float* gamma, const float* beta,
                                     const int32_t* film_row, int32_t film_ld, int32_t off, int32_t width, float* film,
                                     float* stat, void* stream) {
  VNQA_REQUIRE(cond && gamma && beta && film_row && film && stat && off >= 0 && off + width <= N, "vnqa_ln_slice_forward: bad arguments");
  ln_slice_fwd_kernel<<<rows, 256, 0, as_stream(stream)>>>(cond, N, gamma, beta, film_row, film_ld, off, width, film, stat);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_ln_slice_backward(const float* cond, int32_t rows, int32_t N, const float* gamma,
                                      const int32_t* film_row, int32_t film_ld, int32_t off, int32_t width,
                                      const float* dfilm, const float* stat, float* dcond, float* dgamma, float* dbeta,
                                      void* stream) {
  VNQA_REQUIRE(cond && gamma && film_row && dfilm && stat && dcond && dgamma && dbeta && off + width <= N, "vnqa_ln_slice_backward: bad arguments");
  ln_slice_bwd_kernel<<<rows, 256, 0, as_stream(stream)>>>(cond, N, gamma, film_row, film_ld, off, width, dfilm, stat,
                                                          dcond, dgamma, dbeta);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
