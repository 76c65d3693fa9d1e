/*
 * vnqa_b200.h - C ABI of libvnqa_b200.so: hand-written sm_100a kernels for the VideoNavQA FiLM hot path.
 *
 * The reference (catalina17/VideoNavQA) is pure PyTorch: it has no FFI of its own.  The entry points below
 * are what its nn.Module forward/backward for this path bind to, one per op group; each comment cites the
 * reference lines (relative to the reference root) the call replaces.  All pointers are DEVICE pointers;
 * `stream` is a cudaStream_t passed as void*.  No torch types cross this boundary.  Every function returns 0 on
 * success or a negative code (-1 bad argument, -2 CUDA error, -3 tensor-map error); vnqa_last_error() gives the
 * thread-local message.  The library allocates no device memory: callers own every buffer.
 *
 * Activation layout ("frame-image major NHWC"): a batch of clips is flattened to n_img = sum(v_lens) frame
 * images ordered frame-major (all clips of frame 0, then frame 1, ...; clips sorted by descending length as
 * the reference requires, models/film_attn_pt_stem.py:184-186), each image stored as [10][13][C].
 * Element type of every tensor that feeds a tensor-core MMA (activations, packed weights, backward signals) is chosen
 * per call by `act_dtype`: VNQA_ACT_FP16 = IEEE half (the default of the Python layer; kind::f16 MMAs) or
 * VNQA_ACT_TF32 = fp32 words holding tf32-representable values (kind::tf32 MMAs, half the rate).  Both have a 10-bit
 * mantissa; producers round once, to nearest, when they store.  Accumulation, BatchNorm statistics, the small-GEMM
 * tail, gradients of parameters and the optimizer are fp32 in either mode.  `void*` operands are such tensors;
 * `float*` operands are always fp32.
 */
#ifndef VNQA_B200_H
#define VNQA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNQA_ABI_VERSION 3
#define VNQA_IMG_H 10
#define VNQA_IMG_W 13
#define VNQA_IMG_PIX 130 /* hard-coded 130 = 10*13 in models/film_attn_pt_stem.py:56 */

int vnqa_abi_version(void);
const char* vnqa_last_error(void);
/* 0 iff the current CUDA device is sm_100 (B200); the product path has no other backend. */
int vnqa_device_check(void);
int vnqa_sm_count(void);
/* Pipeline bring-up only (libvnqa_b200_dbg.so, built with -DVNQA_DEBUG_MBAR): `host_mapped` is pinned host memory of
 * 148*8 records x 32 uint64 in which a tensor-core kernel whose mbarrier wait exceeds 2 s records who is waiting and
 * the raw state of the CTA's barriers (csrc/ptx.cuh).  The release library accepts the call and ignores the buffer.
 * vnqa_debug_build() returns 1 for the debug library. */
int vnqa_debug_install(void* host_mapped);
int vnqa_debug_build(void);

/* ---- K0: (B, C, 10, 13, T) fp32, T innermost  ->  frame-image-major NHWC [n_img][130][C].
 * Replaces the strided slice v_input[:ct, :, :, :, i] (models/film_attn_pt_stem.py:210).
 * img_of[b*T + t] is the image index of (clip b, frame t), or -1 when t >= v_lens[b]. */
int vnqa_frames_to_nhwc(const float* v_input, void* x_out, const int32_t* img_of, int32_t B, int32_t C, int32_t T,
                        int32_t act_dtype, void* stream);
/* Same result from FRAME-BLOCKED memory: frame t of clip b is one contiguous (C, 10, 13) block at
 * v_input + b*stride_b + t*stride_t (strides in floats).  This is the (B, C, 10, 13, T) VIEW the reference's drivers
 * build from the per-frame stem outputs (torch.stack(...).permute(1, 2, 3, 4, 0)[perm_idx], q_and_v_eval.py:110-116);
 * reading it in place saves a 2 x 596 MB relayout pass. */
int vnqa_frames_blocked_to_nhwc(const float* v_input, void* x_out, const int32_t* img_of, int32_t B, int32_t C, int32_t T,
                                int64_t stride_b, int64_t stride_t, int32_t act_dtype, void* stream);
/* (no backward: v_input needs no gradient - the stem is frozen, eval/q_and_v_eval.py:101-110) */

/* ---- weight packing (derived caches; PyTorch-native layouts stay the source of truth).
 * conv weight (Cout, Cin, kh, kw) -> [taps][rows_pad][Cin] K-major rows (zero rows for cout..rows_pad).
 * transpose_flip != 0 builds the data-gradient operand instead: [taps][rows_pad >= Cin][Cout] with taps mirrored. */
int vnqa_pack_conv_weight(const float* w, void* packed, int32_t cout, int32_t cin, int32_t taps,
                          int32_t rows_pad, int32_t transpose_flip, int32_t act_dtype, void* stream);
/* out[i] = fp16(scale * in[i]): fp32 -> fp16 producers at the edge of the fp16 domain (d feats, packed Linear weights) */
int vnqa_to_half(const float* in, void* out, int64_t n, float scale, void* stream);
/* Linear weight (N, C*130) in NCHW-flatten order (col = c*130 + p, models/film_attn_pt_stem.py:244) ->
 * (n_pad, 130*C) in NHWC order (col = p*C + c); inverse != 0 maps a gradient (n rows) back. */
int vnqa_permute_fc_weight(const float* w, float* out, int32_t n, int32_t n_pad, int32_t C, int32_t inverse,
                           void* stream);

/* ---- data path (eval/dataset.py:57-93 + the .float() of q_and_v_eval.py:93):
 * frames [B][T][H][W][3] uint8 (decoded frames as OpenCV delivers them, already subsampled, zero frames behind each
 * clip's end)  ->  out (B, 3, H, W, T) fp32 with out = lut256[u8]; lut256[i] = (float)((double)i / 255.0) makes it
 * bit-identical to the reference's float64 division and cast. */
int vnqa_frames_u8_to_video(const uint8_t* frames, float* out, const float* lut256, int32_t B, int32_t T, int32_t H,
                            int32_t W, void* stream);

/* ---- K1/K2/K3: implicit-GEMM convolution / GEMM on tcgen05 tensor cores (kind::tf32, fp32 accumulate in TMEM).
 * D[co][pixel] = sum_{tap,ci} Wp[tap][co][ci] * X[pixel + tap][ci], 128 x 208 tiles (208 = 8 images x 2 rows x 13),
 * operands staged by TMA (the zero halo is the TMA out-of-bounds fill), fused epilogues below. */
enum {
  VNQA_EPI_BIAS_RELU = 0,        /* out = relu(acc + bias)                       conv1x1 skip (:219), gp tail (gp :228) */
  VNQA_EPI_BIAS_RELU_AFFINE = 1, /* out = relu(acc + bias) * scale + shift       conv_init + BN eval (:211) */
  VNQA_EPI_BIAS_RELU_STATS = 2,  /* out = relu(acc + bias); per-(frame,ch) sum/sumsq  conv_init + BN train (:211) */
  VNQA_EPI_FILM = 3,             /* z = acc + bias; out = relu(a*z + b) + res    ResBlock body (:224-241) */
  VNQA_EPI_PARTIAL = 4,          /* out[split][row][co] = acc                    split-K GEMM (fc_embed_attn :244) */
  VNQA_EPI_DGRAD_RES = 5,        /* out = (acc + add) * [mask > 0]               dX of conv3x3 + skip, through ReLU (:219) */
  VNQA_EPI_PLAIN = 7,            /* out = acc                                    dX of conv1x1 / Linear */
  VNQA_EPI_BIAS = 8,             /* out = acc + bias                             frozen stem conv11/21/31 (obj_detector.py:71,76,81) */
  VNQA_EPI_BIAS_AFFINE_RELU = 9  /* out = [maxpool2x2] relu((acc + bias) * scale + shift)   stem conv12/22/32 + BN eval + ReLU
                                    [+ pool] (obj_detector.py:71-84) */
};

/* Operand / activation storage of the tensor-core kernels ("act" below): tf32 = fp32 storage holding tf32-representable
 * values (kind::tf32 MMA); fp16 = half storage with the same 10-bit mantissa (kind::f16 MMA: twice the rate, half the
 * bytes).  Accumulation is fp32 either way.  With fp16 the backward signals are stored multiplied by a power-of-two
 * loss scale and every kernel that leaves the fp16 domain multiplies by grad_unscale = 1 / scale. */
enum { VNQA_ACT_TF32 = 0, VNQA_ACT_FP16 = 1 };

typedef struct {
  /* operands ("act" = float or __half, see act_dtype) */
  const void* w_packed;  /* act [taps][rows_pad][cin] from vnqa_pack_conv_weight (flat: [rows_pad][cin*splits]) */
  const void* x;         /* act activations [n_img][img_h*img_w][cin] (conv) or [n_img][x_ld] (flat) */
  int32_t taps;          /* 1 or 9 */
  int32_t cin;           /* reduction length per tap (per split when flat); tf32: multiple of 32, fp16: of 8 */
  int32_t cout;          /* real output channels */
  int32_t rows_pad;      /* cout rounded up to 128 */
  int32_t n_img;         /* conv: images; flat: rows of x */
  int32_t flat;          /* 0: conv over img_h x img_w images (10 x 13 by default); 1: plain GEMM */
  int32_t x_ld;          /* flat only: row stride of x in elements (>= cin*splits) */
  int32_t splits;        /* flat only: split-K factor */
  int32_t epi;
  int32_t no_round;      /* tf32 only, != 0: store full fp32 (output does not feed an MMA) */
  /* epilogue operands (unused ones may be NULL) */
  void* out;             /* act [rows][cout]; always fp32 for STATS and PARTIAL ([splits][n_img][cout]) and with out_f32 */
  const float* bias;     /* [cout] */
  const float* scale;    /* AFFINE: [cout] */
  const float* shift;    /* AFFINE: [cout] */
  const int32_t* img_frame; /* STATS: frame of each image */
  double* stat_sum;      /* STATS: [T][cout], atomically accumulated */
  double* stat_sq;       /* STATS: [T][cout] */
  const float* film;     /* FILM: film values, row r at film + r*film_ld */
  const int32_t* film_row; /* FILM: row of `film` for each image (= frame*B + clip) */
  int32_t film_ld;
  int32_t alpha_off;     /* FILM: alpha at [alpha_off + co], beta at [alpha_off + cout + co] */
  const void* res;       /* act. FILM: residual input [rows][cout]; DGRAD_RES: ReLU mask source */
  void* z_out;           /* act. FILM: optional pre-FiLM output z (training) */
  const void* add;       /* act. DGRAD_RES: tensor added before masking (skip-path gradient) */
  float* dbias;          /* DGRAD_RES: optional; += grad_unscale * per-channel sum of the written gradient (atomic) */
  int32_t act_dtype;     /* VNQA_ACT_TF32 / VNQA_ACT_FP16 */
  int32_t out_f32;       /* fp16 only, BIAS_RELU / PLAIN: store fp32 (tensors that leave the fp16 domain) */
  float grad_unscale;    /* 0 is read as 1 */
  /* image geometry of the conv (flat == 0); 0 = the FiLM trunk's 10 x 13.  img_w must be 13, 26 or 52 (a 208-pixel
   * tile is then (8, 4, 2) images x 2 rows x img_w), img_h even: the frozen stem's 40 x 52 / 20 x 26 / 10 x 13 maps */
  int32_t img_h;
  int32_t img_w;
  int32_t pool;          /* BIAS_AFFINE_RELU, img_w 26 / 52: write the 2x2 max-pooled map [n_img][img_h/2 * img_w/2][cout] */
} vnqa_conv_args;

int vnqa_conv_gemm(const vnqa_conv_args* args, void* stream);

/* ---- frozen stem (models/obj_detector.py:69-86, run under no_grad in eval mode by
 * eval/q_and_v_eval.py:104-110).  Its six 3x3 convolutions are vnqa_conv_gemm with img_h / img_w / pool and the BIAS /
 * BIAS_AFFINE_RELU epilogues; these two calls are the layout changes at its ends.
 * x (n_img, C, P) fp32 NCHW -> out [n_img][P][C] act, out = x * scale[c] + shift[c] (bn_input in eval mode, :70;
 * scale == shift == NULL: plain relayout). */
int vnqa_nchw_to_nhwc_affine(const float* x, void* out, const float* scale, const float* shift, int32_t n_img,
                             int32_t C, int32_t P, int32_t act_dtype, void* stream);
/* x [n_img][P][C] (fp32, or half when x_half != 0) -> out (n_img, C, P) fp32 NCHW, the layout the reference's callers
 * stack per frame (q_and_v_eval.py:110) */
int vnqa_nhwc_to_nchw(const void* x, float* out, int32_t n_img, int32_t C, int32_t P, int32_t x_half, void* stream);

/* ---- K4: weight gradient dW[co][ci][tap] = grad_unscale * sum_rows dY[row][co] * X[row (+) tap][ci] on tcgen05
 * (MN-major operands of type act_dtype, split-K over pixels, DETERMINISTIC turn-ordered fp32 reduction: two launches on
 * the same operands give the same bits).  dw is OVERWRITTEN in PyTorch's (Cout, Cin, kh, kw) layout.
 * ws: workspace of vnqa_conv_wgrad_ws_bytes() bytes whose leading *zero_bytes bytes (turn counters) must be zero on
 * entry; the kernel leaves them zero again.  flat != 0: plain dY^T X with dy [n_img][cout], x [n_img][cin]. */
int64_t vnqa_conv_wgrad_ws_bytes(int32_t cout, int32_t cin, int32_t taps, int64_t* zero_bytes);
int vnqa_conv_wgrad(const void* dy, const void* x, float* dw, void* ws, int32_t n_img, int32_t cout, int32_t cin,
                    int32_t taps, int32_t flat, int32_t act_dtype, float grad_unscale, void* stream);

/* ---- small fp32 GEMM (FFMA, exact fp32): C = [C +] A B^T + bias, optional ReLU.  A(m,k) = a[m*a_rs + k*a_cs],
 * B(n,k) = b[n*b_rs + k*b_cs].  ksplit > 1 splits K over gridDim.z with atomic accumulation into C (C must be
 * zero-initialised or hold the value to accumulate onto; relu/bias are then not allowed).
 * Linears of :179 (FiLM generator), :301 (classifier) and their gradients. */
int vnqa_sgemm(const float* a, const float* b, float* c, const float* bias, int32_t M, int32_t N, int32_t K,
               int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t ldc, int32_t relu, int32_t accumulate,
               int32_t ksplit, void* stream);
/* out[n] (+)= sum_rows x[row][n] */
int vnqa_colsum(const float* x, int32_t rows, int32_t n, int64_t ld, float* out, int32_t accumulate, void* stream);

/* ---- BatchNorm2d after ReLU, one call per processed frame in the reference (:211, quirk Q6) */
int vnqa_bn_finalize_train(const double* stat_sum, const double* stat_sq, const int32_t* frame_ct, int32_t n_frames,
                           int32_t C, const float* gamma, const float* beta, float* running_mean, float* running_var,
                           float momentum, float eps, float* scale, float* shift, float* mean, float* rstd,
                           void* stream);
int vnqa_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float eps, int32_t C, float* scale, float* shift, void* stream);
/* x = y * scale[frame] + shift[frame]   (scale/shift are [T][C]) */
int vnqa_bn_apply(const float* y, void* x, const int32_t* img_frame, const float* scale, const float* shift,
                  int32_t n_img, int32_t C, int32_t act_dtype, void* stream);
/* dy = BN-backward(dx) masked by the preceding ReLU (y > 0); dbias += sum dy; dgamma/dbeta accumulated. */
int vnqa_bn_backward(const void* dx, const float* y, const int32_t* img_frame, const int32_t* frame_ct,
                     int32_t n_frames, const float* mean, const float* rstd, const float* gamma, int32_t n_img,
                     int32_t C, float* s1, float* s2, void* dy, float* dbias, float* dgamma, float* dbeta,
                     int32_t act_dtype, float grad_unscale, void* stream);

/* ---- FiLM backward (autograd of :229-239): dz = alpha * dout * [alpha z + beta > 0]; dfilm rows get
 * dalpha/dbeta; mask_relu != 0 additionally applies the gradient of the FiLM generator's final ReLU (:179), i.e.
 * zero where alpha/beta == 0 (attention / global-pooling models; the multi-hop model has no such ReLU);
 * dbias += sum dz. */
int vnqa_film_backward(const void* dout, const void* z, const float* film, const int32_t* film_row, int32_t film_ld,
                       int32_t alpha_off, int32_t n_img, int32_t C, void* dz, float* dfilm, float* dbias,
                       int32_t mask_relu, int32_t act_dtype, float grad_unscale, void* stream);

/* out[dst_row[r]][n] = bias[n] + sum_s partial[s][r][n] */
int vnqa_splitk_reduce(const float* partial, int32_t splits, int32_t rows, int32_t N, const float* bias,
                       const int32_t* dst_row, float* out, void* stream);
/* out[r][:] = src[src_row[r]][:]  (round_tf32 != 0 when the result feeds an MMA) */
int vnqa_gather_rows(const float* src, const int32_t* src_row, int32_t rows, int32_t N, float* out,
                     int32_t round_tf32, void* stream);

/* ---- global temporal max pooling (models/film_global_pooling_pt_stem.py:230-235) */
int vnqa_gp_maxpool(const float* y, const int32_t* img_of, int32_t B, int32_t T, int32_t tail, float* pooled,
                    int32_t* arg_img, void* stream);
int vnqa_gp_maxpool_backward(const float* dpooled, const float* pooled, const int32_t* arg_img, int32_t B, int32_t tail,
                             float* dy, void* stream);

/* ---- K6: question encoder (compute_film_values, :144-171): gx = W_ih embed(q) + b_ih + b_hh once, then the
 * packed LSTM chained over n_pass frames with carried state (quirk Q3). */
int vnqa_embed_gates(const int64_t* q_input, const int32_t* q_lens, const float* embed, const float* w_ih,
                     const float* b_ih, const float* b_hh, int32_t B, int32_t L, int32_t E, int32_t H, float* gx,
                     void* stream);
/* out_states (optional, [n_pass][B][L][H], zero-initialised): every step's output = pad_packed_sequence view used by
 * the multi-hop model (models/time_multi_hop_pt_stem.py:136-142). */
int vnqa_lstm_chain_forward(const float* gx, const int32_t* q_lens, const float* whh_t, float* h_state, float* c_state,
                            int32_t n_pass, int32_t B, int32_t L, int32_t H, float* h_pass, float* save_gates,
                            float* save_c, float* save_hprev, float* out_states, void* stream);
/* dh_pass ([n_pass][B][H]) and/or dh_steps ([n_pass][B][L][H]): incoming gradients; at least one is required. */
int vnqa_lstm_chain_backward(const float* dh_pass, const int32_t* q_lens, const float* whh, const float* save_gates,
                             const float* save_c, const float* c0, int32_t n_pass, int32_t B, int32_t L, int32_t H,
                             float* dgates, float* dgx, const float* dh_steps, void* stream);
/* dembed[q[b][t]][:] += dx[b][t][:] for t < q_len[b]  (dembed pre-initialised by the caller) */
int vnqa_embed_backward(const int64_t* q_input, const int32_t* q_lens, const float* dx, int32_t B, int32_t L, int32_t E,
                        float* dembed, void* stream);
int vnqa_embed_gather(const int64_t* q_input, const int32_t* q_lens, const float* embed, int32_t B, int32_t L,
                      int32_t E, float* x, void* stream);

/* ---- K9: multi-hop FiLM decoder of TimeMultiHopFiLMPretrainedStem (models/time_multi_hop_pt_stem.py:124-184):
 * encoder LayerNorm + NB chained hops of unmasked word attention per frame image; then per hop a Linear (vnqa_sgemm)
 * and a full-width LayerNorm of which only the block's 2C-wide slice is written to the FiLM matrix. */
int vnqa_tmh_hops_forward(const float* states, const float* h_last, const int32_t* img_frame, const int32_t* img_clip,
                          int32_t n_img, int32_t B, int32_t L, int32_t H, int32_t W, int32_t NB, const float* enc_g,
                          const float* enc_b, const float* w_att, const float* b_att, float* hvec, float* enc_out,
                          float* enc_stat, float* coefs, void* stream);
int vnqa_tmh_hops_backward(const float* states, const float* h_last, const int32_t* img_frame, const int32_t* img_clip,
                           const int32_t* q_lens, int32_t n_img, int32_t B, int32_t L, int32_t H, int32_t W, int32_t NB,
                           const float* enc_g, const float* w_att, const float* hvec, const float* enc_out,
                           const float* enc_stat, const float* coefs, const float* dhvec, float* dstates, float* dw_att,
                           float* db_att, float* denc_g, float* denc_b, void* stream);
int vnqa_ln_slice_forward(const float* cond, int32_t rows, int32_t N, const float* gamma, const float* beta,
                          const int32_t* film_row, int32_t film_ld, int32_t off, int32_t width, float* film, float* stat,
                          void* stream);
int vnqa_ln_slice_backward(const float* cond, int32_t rows, int32_t N, const float* gamma, const int32_t* film_row,
                           int32_t film_ld, int32_t off, int32_t width, const float* dfilm, const float* stat,
                           float* dcond, float* dgamma, float* dbeta, void* stream);

/* ---- K7: temporal attention read-out (:253-295) and its backward */
int vnqa_attn_tail_forward(const float* feats, const int32_t* v_lens, int32_t n_proc, int32_t B, int32_t T, int32_t A,
                           const float* w1, const float* b1, const float* wh, const float* bh, const float* wih_t,
                           const float* whh_t, const float* b_ih, const float* b_hh, float* hs, float* coefs,
                           float* ctxt, float* gates, float* cs, void* stream);
int vnqa_attn_tail_backward(const float* feats, const int32_t* v_lens, int32_t n_proc, int32_t B, int32_t T, int32_t A,
                            const float* w1, const float* wh, const float* wih, const float* whh, const float* hs,
                            const float* coefs, const float* gates, const float* cs, const float* dhs, float* dfeats,
                            float* dgates, float* dw1, float* db1, float* dwh, float* dbh, void* stream);

/* ---- global-norm clip + Adam over FLAT fp32 buffers (n multiple of 4, 16-byte aligned);
 * replaces clip_grad_norm(model.parameters(), 1.0) + optimizer.step() of eval/q_and_v_eval.py:137-139.
 * vnqa_sumsq_partials writes vnqa_opt_num_partials() per-block sums of g^2 (fp64, no atomics -> deterministic);
 * vnqa_clip_adam reduces them in a fixed order, scales the gradient by grad_scale * min(1, max_norm / (norm + 1e-6))
 * (max_norm <= 0: no clipping, partials may be NULL) and applies torch.optim.Adam's update (no amsgrad, no weight
 * decay) for step number `step` (1-based; bias corrections and 1 - beta formed in double on the host like PyTorch).  norm_out (optional) receives
 * the total gradient norm that clip_grad_norm_ would return. */
int vnqa_opt_num_partials(void);
int vnqa_sumsq_partials(const float* g, int64_t n, double* partials, void* stream);
int vnqa_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const double* partials, int32_t n_partials,
                   float grad_scale, float max_norm, double lr, double beta1, double beta2, double eps, int64_t step,
                   float* norm_out, void* stream);
/* Same update with the step count kept ON THE DEVICE (*step_dev is incremented, then used): nothing in the launch
 * depends on a host-side counter, so a captured CUDA graph of the train step can be replayed. */
int vnqa_clip_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const double* partials,
                       int32_t n_partials, float grad_scale, float max_norm, double lr, double beta1, double beta2,
                       double eps, int64_t* step_dev, float* norm_out, void* stream);

/* ---- K10 (new - the reference is single-GPU): gradient all-reduce over NVLink / NVSwitch.
 * One process per GPU owns one ncclComm_t and one dedicated communication stream inside the library.  NCCL is
 * dlopen'ed (vnqa_nccl_load; NULL = "libnccl.so.2").  Rank 0 creates the 128-byte unique id, every rank calls
 * vnqa_nccl_init with it (collective).  vnqa_nccl_allreduce_sum_f32 enqueues an IN-PLACE SUM all-reduce on the
 * communication stream ordered after everything queued so far on `after_stream` and returns at once, so the
 * collective overlaps the remaining backward kernels; vnqa_nccl_join makes a stream wait for all queued collectives.
 * SUM, not mean: the reference loss is CrossEntropyLoss(reduction='sum') (eval.sh:15).  Errors return -4. */
int vnqa_nccl_load(const char* libpath);
int vnqa_nccl_version(void);
int vnqa_nccl_unique_id(void* out128);
int vnqa_nccl_init(const void* id128, int32_t rank, int32_t nranks);
int vnqa_nccl_rank(void);
int vnqa_nccl_nranks(void);
void* vnqa_nccl_stream(void);
int vnqa_nccl_allreduce_sum_f32(float* buf, int64_t count, void* after_stream);
int vnqa_nccl_allgather_f64(const double* send, double* recv, int64_t count, void* after_stream);
int vnqa_nccl_broadcast_f32(float* buf, int64_t count, int32_t root, void* after_stream);
int vnqa_nccl_join(void* stream);
int vnqa_nccl_destroy(void);
int vnqa_nccl_abort(void);

#ifdef __cplusplus
}
#endif
#endif
