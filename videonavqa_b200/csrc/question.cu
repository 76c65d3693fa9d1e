// K6: question encoder - embedding gather + input-gate pre-GEMM (once) and the chained packed LSTM
// (state carried across frames).  Replaces compute_film_values of
// models/film_attn_pt_stem.py:144-171 (the Linear+ReLU of :179 is vnqa_sgemm).
//
// The recurrence is latency-bound: n_pass * q_len sequential 4H x H mat-vecs per clip.  One CTA per clip with
// 4H threads; thread j owns gate row j.  W_hh^T lives on chip for the whole chain: rows k < KS in shared memory
// (conflict-free [k][4H] layout), the remaining H-KS rows in registers.
#include "common.h"

namespace vnqa {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gx[b][t][j] = b_ih[j] + b_hh[j] + sum_e W_ih[j][e] * embed[q[b][t]][e]   for t < q_len[b]
__global__ void __launch_bounds__(128) embed_gates_kernel(const long long* __restrict__ q, const int* __restrict__ q_lens,
                                                          const float* __restrict__ embed, const float* __restrict__ w_ih,
                                                          const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                          int L, int E, int H4, float* __restrict__ gx) {
  extern __shared__ float xs[];
  const int b = blockIdx.y, t = blockIdx.x;
  if (t >= q_lens[b]) return;
  const long long tok = q[(long long)b * L + t];
  for (int e = threadIdx.x; e < E; e += blockDim.x) xs[e] = embed[tok * E + e];
  __syncthreads();
  for (int j = threadIdx.x; j < H4; j += blockDim.x) {
    const float* wr = w_ih + (long long)j * E;
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(__ldg(wr + e), xs[e], acc);
    gx[((long long)b * L + t) * H4 + j] = acc + b_ih[j] + b_hh[j];
  }
}

// x[b][t][:] = embed[q[b][t]][:] (zero for t >= q_len[b]);  dembed[q[b][t]][:] += dx[b][t][:]
__global__ void embed_gather_kernel(const long long* __restrict__ q, const int* __restrict__ q_lens,
                                    const float* __restrict__ embed, int L, int E, float* __restrict__ x) {
  const int b = blockIdx.y, t = blockIdx.x;
  const bool valid = t < q_lens[b];
  const long long tok = valid ? q[(long long)b * L + t] : 0;
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    x[((long long)b * L + t) * E + e] = valid ? embed[tok * E + e] : 0.f;
}
__global__ void embed_backward_kernel(const long long* __restrict__ q, const int* __restrict__ q_lens,
                                      const float* __restrict__ dx, int L, int E, float* __restrict__ dembed) {
  const int b = blockIdx.y, t = blockIdx.x;
  if (t >= q_lens[b]) return;
  const long long tok = q[(long long)b * L + t];
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    atomicAdd(dembed + tok * E + e, dx[((long long)b * L + t) * E + e]);
}

// Forward chain.  whh_t = W_hh^T [H][4H].  State (h, c) [B][H] is read and written (carried across calls unless
// the caller zeroes it = init_hidden()).  h_pass[p][b][:] = h after pass p.  When save_* are non-null every step
// stores its activated gates / c / previous h for BPTT: index ((p*B + b)*L + t).
template <int H, int KS>
__global__ void __launch_bounds__(4 * H, 1)
lstm_chain_fwd_kernel(const float* __restrict__ gx, const int* __restrict__ q_lens, const float* __restrict__ whh_t,
                      float* __restrict__ h_state, float* __restrict__ c_state, int n_pass, int B, int L,
                      float* __restrict__ h_pass, float* __restrict__ save_gates, float* __restrict__ save_c,
                      float* __restrict__ save_h, float* __restrict__ out_states) {
  constexpr int H4 = 4 * H;
  constexpr int KR = H - KS;  // rows kept in registers
  extern __shared__ float sm[];
  float* wsm = sm;                 // [KS][4H]
  float* hs = sm + KS * H4;        // [H]
  float* gs = hs + H;              // [4H]
  const int b = blockIdx.x, j = threadIdx.x;
  for (int i = j; i < KS * H4; i += H4) wsm[i] = whh_t[i];
  float wreg[KR > 0 ? KR : 1];
#pragma unroll
  for (int k = 0; k < KR; ++k) wreg[k] = whh_t[(KS + k) * H4 + j];
  float c = 0.f;
  if (j < H) {
    hs[j] = h_state[(long long)b * H + j];
    c = c_state[(long long)b * H + j];
  }
  __syncthreads();
  const int ql = q_lens[b];
  for (int p = 0; p < n_pass; ++p) {
    for (int t = 0; t < ql; ++t) {
      float acc = gx[((long long)b * L + t) * H4 + j];
#pragma unroll 8
      for (int k = 0; k < KS; ++k) acc = fmaf(wsm[k * H4 + j], hs[k], acc);
#pragma unroll
      for (int k = 0; k < KR; ++k) acc = fmaf(wreg[k], hs[KS + k], acc);
      gs[j] = acc;
      __syncthreads();
      if (j < H) {
        const float ig = sigmoidf_(gs[j]), fg = sigmoidf_(gs[H + j]), gg = tanhf(gs[2 * H + j]),
                    og = sigmoidf_(gs[3 * H + j]);
        c = fg * c + ig * gg;
        const float h = og * tanhf(c);
        const float h_prev = hs[j];
        hs[j] = h;
        if (out_states) out_states[(((long long)p * B + b) * L + t) * H + j] = h;   // pad_packed_sequence view
        if (save_gates) {
          const long long s = ((long long)p * B + b) * L + t;
          save_gates[s * H4 + j] = ig;
          save_gates[s * H4 + H + j] = fg;
          save_gates[s * H4 + 2 * H + j] = gg;
          save_gates[s * H4 + 3 * H + j] = og;
          save_c[s * H + j] = c;
          save_h[s * H + j] = h_prev;  // input state of this step (dW_hh = sum dG (x) h_prev)
        }
      }
      __syncthreads();
    }
    if (j < H) h_pass[((long long)p * B + b) * H + j] = hs[j];
  }
  if (j < H) {
    h_state[(long long)b * H + j] = hs[j];
    c_state[(long long)b * H + j] = c;
  }
}

// BPTT through the whole chain.  dh_pass[p][b][:] is the gradient arriving at the state after pass p.
// Writes the pre-activation gate gradients of every step to dgates ((p*B+b)*L+t)[4H] (for dW_hh = dG^T h_prev by
// GEMM) and accumulates them over passes into dgx[b][t][4H] (for dW_ih / dEmbed / db).  whh = W_hh [4H][H].
template <int H, int KS>
__global__ void __launch_bounds__(4 * H, 1)
lstm_chain_bwd_kernel(const float* __restrict__ dh_pass, const int* __restrict__ q_lens, const float* __restrict__ whh,
                      const float* __restrict__ save_gates, const float* __restrict__ save_c,
                      const float* __restrict__ c0, int n_pass, int B, int L, float* __restrict__ dgates,
                      float* __restrict__ dgx, const float* __restrict__ dh_steps) {
  // dh_steps (optional, [n_pass][B][L][H]): gradient arriving at the OUTPUT of every step (multi-hop model, where
  // all word states are consumed); dh_pass (optional): gradient at the state after each pass.
  constexpr int H4 = 4 * H;
  constexpr int KR = H - KS;
  extern __shared__ float sm[];
  float* wsm = sm;                // [KS][4H] : rows jj < KS of each quarter, see below
  float* dgs = sm + KS * H4;      // [4H] gate gradients of the current step
  float* part = dgs + H4;         // [4][H] partial dh sums
  const int b = blockIdx.x, tid = threadIdx.x;
  const int qd = tid / H, k = tid % H;  // thread (qd, k) sums rows j = qd*H + jj of W_hh, column k
  // wsm[jj][qd*H + k] = W_hh[qd*H + jj][k]  -> reads wsm[jj*4H + tid], conflict-free
  for (int i = tid; i < KS * H4; i += H4) {
    const int jj = i / H4, col = i % H4, q2 = col / H, k2 = col % H;
    wsm[i] = whh[(long long)(q2 * H + jj) * H + k2];
  }
  float wreg[KR > 0 ? KR : 1];
#pragma unroll
  for (int r = 0; r < KR; ++r) wreg[r] = whh[(long long)(qd * H + KS + r) * H + k];
  for (int i = tid; i < L * H4; i += H4) dgx[(long long)b * L * H4 + i] = 0.f;
  __syncthreads();
  const int ql = q_lens[b];
  float dh = 0.f, dc = 0.f;  // threads tid < H hold the running state gradients
  for (int p = n_pass - 1; p >= 0; --p) {
    if (tid < H && dh_pass) dh += dh_pass[((long long)p * B + b) * H + tid];
    for (int t = ql - 1; t >= 0; --t) {
      const long long s = ((long long)p * B + b) * L + t;
      if (tid < H) {
        if (dh_steps) dh += dh_steps[s * H + tid];
        const float ig = save_gates[s * H4 + tid], fg = save_gates[s * H4 + H + tid],
                    gg = save_gates[s * H4 + 2 * H + tid], og = save_gates[s * H4 + 3 * H + tid];
        const float cc = save_c[s * H + tid];
        float cprev;
        if (t > 0) cprev = save_c[(s - 1) * H + tid];
        else if (p > 0) cprev = save_c[((((long long)(p - 1)) * B + b) * L + (ql - 1)) * H + tid];
        else cprev = c0[(long long)b * H + tid];
        const float tc = tanhf(cc);
        const float dct = dc + dh * og * (1.f - tc * tc);
        const float d_i = dct * gg * ig * (1.f - ig);
        const float d_f = dct * cprev * fg * (1.f - fg);
        const float d_g = dct * ig * (1.f - gg * gg);
        const float d_o = dh * tc * og * (1.f - og);
        dc = dct * fg;
        dgs[tid] = d_i; dgs[H + tid] = d_f; dgs[2 * H + tid] = d_g; dgs[3 * H + tid] = d_o;
      }
      __syncthreads();
      {
        const float g = dgs[tid];
        dgates[s * H4 + tid] = g;
        dgx[((long long)b * L + t) * H4 + tid] += g;
      }
      float acc = 0.f;
#pragma unroll 8
      for (int jj = 0; jj < KS; ++jj) acc = fmaf(wsm[jj * H4 + tid], dgs[qd * H + jj], acc);
#pragma unroll
      for (int r = 0; r < KR; ++r) acc = fmaf(wreg[r], dgs[qd * H + KS + r], acc);
      part[tid] = acc;
      __syncthreads();
      if (tid < H) dh = part[tid] + part[H + tid] + part[2 * H + tid] + part[3 * H + tid];
      __syncthreads();
    }
  }
}

template <int H, int KS>
static int launch_fwd(const float* gx, const int* q_lens, const float* whh_t, float* h_state, float* c_state, int n_pass,
                      int B, int L, float* h_pass, float* sg, float* sc, float* sh, float* states, cudaStream_t st) {
  auto kern = lstm_chain_fwd_kernel<H, KS>;
  const int smem = (KS * 4 * H + H + 4 * H) * 4;
  VNQA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<B, 4 * H, smem, st>>>(gx, q_lens, whh_t, h_state, c_state, n_pass, B, L, h_pass, sg, sc, sh, states);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
template <int H, int KS>
static int launch_bwd(const float* dh_pass, const int* q_lens, const float* whh, const float* sg, const float* sc,
                      const float* c0, int n_pass, int B, int L, float* dgates, float* dgx, const float* dh_steps,
                      cudaStream_t st) {
  auto kern = lstm_chain_bwd_kernel<H, KS>;
  const int smem = (KS * 4 * H + 4 * H + 4 * H) * 4;
  VNQA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<B, 4 * H, smem, st>>>(dh_pass, q_lens, whh, sg, sc, c0, n_pass, B, L, dgates, dgx, dh_steps);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_embed_gates(const int64_t* q_input, const int32_t* q_lens, const float* embed, const float* w_ih,
                                const float* b_ih, const float* b_hh, int32_t B, int32_t L, int32_t E, int32_t H,
                                float* gx, void* stream) {
  VNQA_REQUIRE(q_input && q_lens && embed && w_ih && b_ih && b_hh && gx, "vnqa_embed_gates: null pointer");
  dim3 grid(L, B);
  embed_gates_kernel<<<grid, 128, E * 4, as_stream(stream)>>>((const long long*)q_input, q_lens, embed, w_ih, b_ih, b_hh,
                                                              L, E, 4 * H, gx);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_embed_gather(const int64_t* q_input, const int32_t* q_lens, const float* embed, int32_t B, int32_t L,
                                 int32_t E, float* x, void* stream) {
  dim3 grid(L, B);
  embed_gather_kernel<<<grid, 128, 0, as_stream(stream)>>>((const long long*)q_input, q_lens, embed, L, E, x);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_embed_backward(const int64_t* q_input, const int32_t* q_lens, const float* dx, int32_t B, int32_t L,
                                   int32_t E, float* dembed, void* stream) {
  dim3 grid(L, B);
  embed_backward_kernel<<<grid, 128, 0, as_stream(stream)>>>((const long long*)q_input, q_lens, dx, L, E, dembed);
  VNQA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vnqa_lstm_chain_forward(const float* gx, const int32_t* q_lens, const float* whh_t, float* h_state,
                                       float* c_state, int32_t n_pass, int32_t B, int32_t L, int32_t H, float* h_pass,
                                       float* save_gates, float* save_c, float* save_h, float* out_states, void* stream) {
  VNQA_REQUIRE(gx && q_lens && whh_t && h_state && c_state && h_pass, "vnqa_lstm_chain_forward: null pointer");
  cudaStream_t st = as_stream(stream);
  switch (H) {
    case 32: return launch_fwd<32, 32>(gx, q_lens, whh_t, h_state, c_state, n_pass, B, L, h_pass, save_gates, save_c, save_h, out_states, st);
    case 64: return launch_fwd<64, 64>(gx, q_lens, whh_t, h_state, c_state, n_pass, B, L, h_pass, save_gates, save_c, save_h, out_states, st);
    case 128: return launch_fwd<128, 96>(gx, q_lens, whh_t, h_state, c_state, n_pass, B, L, h_pass, save_gates, save_c, save_h, out_states, st);
    default: set_error("vnqa_lstm_chain_forward: hidden size %d not supported (32, 64, 128)", H); return -1;
  }
}

extern "C" int vnqa_lstm_chain_backward(const float* dh_pass, const int32_t* q_lens, const float* whh,
                                        const float* save_gates, const float* save_c, const float* c0, int32_t n_pass,
                                        int32_t B, int32_t L, int32_t H, float* dgates, float* dgx, const float* dh_steps,
                                        void* stream) {
  VNQA_REQUIRE((dh_pass || dh_steps) && q_lens && whh && save_gates && save_c && c0 && dgates && dgx, "vnqa_lstm_chain_backward: null");
  cudaStream_t st = as_stream(stream);
  switch (H) {
    case 32: return launch_bwd<32, 32>(dh_pass, q_lens, whh, save_gates, save_c, c0, n_pass, B, L, dgates, dgx, dh_steps, st);
    case 64: return launch_bwd<64, 64>(dh_pass, q_lens, whh, save_gates, save_c, c0, n_pass, B, L, dgates, dgx, dh_steps, st);
    case 128: return launch_bwd<128, 96>(dh_pass, q_lens, whh, save_gates, save_c, c0, n_pass, B, L, dgates, dgx, dh_steps, st);
    default: set_error("vnqa_lstm_chain_backward: hidden size %d not supported (32, 64, 128)", H); return -1;
  }
}
