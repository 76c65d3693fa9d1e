/* CANDIDATE C ABI for fp16-operand variants of the two tensor-core kernels (round-2 work item).
 *
 * NOT part of libvnqa_b200.so, NOT loaded by the package, NOT YET RUN ON HARDWARE.  Rationale and the CPU precision
 * study behind it: experimental/f16/README.md.  Same conventions as include/vnqa_b200.h (extern "C", raw device
 * pointers, stream as void*, 0 = success, message through vnqa_last_error()).
 *
 * fp16 has the same 10-bit mantissa as tf32, so products are exactly as precise as today's kind::tf32 path, but
 * tcgen05 issues kind::f16 at twice the tf32 rate and every activation tensor is half the HBM bytes.  Accumulation
 * stays fp32 in TMEM.  Backward signals are stored multiplied by a power-of-two loss scale (they reach 1e-7) and the
 * kernels that leave the fp16 domain (weight / bias gradients) multiply by `grad_unscale` = 1 / scale.
 */
#ifndef VNQA_F16_H
#define VNQA_F16_H
#include <stdint.h>

#include "../../include/vnqa_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of vnqa_conv_args (include/vnqa_b200.h) with fp16 tensors where the data feed or leave an MMA. */
typedef struct {
  const void* w_packed;      /* fp16 [taps][rows_pad][cin] from vnqa_pack_conv_weight_h (flat: [rows_pad][cin*splits]) */
  const void* x;             /* fp16 [n_img][130][cin]  (flat: [n_img rows][x_ld]) */
  int32_t taps, cin, cout, rows_pad, n_img, flat, x_ld, splits, epi;
  void* out;                 /* fp16 [rows][cout]; fp32 for VNQA_EPI_BIAS_RELU_STATS, VNQA_EPI_PARTIAL or out_f32 */
  const float* bias;
  const float* scale;
  const float* shift;
  const int32_t* img_frame;
  double* stat_sum;
  double* stat_sq;
  const float* film;
  const int32_t* film_row;
  int32_t film_ld, alpha_off;
  const void* res;           /* fp16: FILM residual / DGRAD_RES ReLU mask */
  void* z_out;               /* fp16, optional (FILM): z = acc + bias saved for backward */
  const void* add;           /* fp16 (DGRAD_RES) */
  float* dbias;              /* fp32, optional (DGRAD_RES): += grad_unscale * column sums */
  float grad_unscale;        /* 0 is read as 1 */
  int32_t out_f32;           /* BIAS_RELU / PLAIN: store fp32 (tails that leave the fp16 domain) */
} vnqa_conv_args_h;

/* conv weight (Cout, Cin, kh, kw) fp32 -> fp16 [taps][rows_pad][Cin] (or the mirrored / transposed data-gradient
 * operand), same meaning as vnqa_pack_conv_weight. */
int vnqa_pack_conv_weight_h(const float* w, void* packed, int32_t cout, int32_t cin, int32_t taps, int32_t rows_pad,
                            int32_t transpose_flip, void* stream);
/* out[i] = fp16(scale * in[i])  (fp32 -> fp16 producers at the edge of the fp16 domain, e.g. d feats) */
int vnqa_to_half(const float* in, void* out, int64_t n, float scale, void* stream);

/* Implicit-GEMM convolution / flat GEMM, tcgen05 kind::f16, fp32 accumulate (replaces vnqa_conv_gemm). */
int vnqa_conv_gemm_h(const vnqa_conv_args_h* args, void* stream);

/* Weight gradient (replaces vnqa_conv_wgrad): dy fp16 [rows][cout], x fp16 [rows][cin]; dw / ws fp32 as before,
 * every partial tile is multiplied by grad_unscale before it is reduced. */
int vnqa_conv_wgrad_h(const void* dy, const void* x, float* dw, float* ws, int32_t n_img, int32_t cout, int32_t cin,
                      int32_t taps, int32_t flat, float grad_unscale, void* stream);

/* ---- HBM-bound kernels with fp16 trunk activations (same meaning as their fp32 namesakes in vnqa_b200.h) ---- */
/* K0: (B, C, 10, 13, T) fp32 -> fp16 [n_img][130][C]; C multiple of 32 */
int vnqa_frames_to_nhwc_h(const float* v_input, void* x_out, const int32_t* img_of, int32_t B, int32_t C, int32_t T,
                          void* stream);
/* x = fp16(y * scale[frame] + shift[frame]); y fp32 (BatchNorm input stays fp32) */
int vnqa_bn_apply_h(const float* y, void* x, const int32_t* img_frame, const float* scale, const float* shift,
                    int32_t n_img, int32_t C, void* stream);
/* dx fp16 (loss-scaled), y fp32 -> dy fp16 (loss-scaled); dbias / dgamma / dbeta += grad_unscale * sums */
int vnqa_bn_backward_h(const void* dx, const float* y, const int32_t* img_frame, const int32_t* frame_ct,
                       int32_t n_frames, const float* mean, const float* rstd, const float* gamma, int32_t n_img,
                       int32_t C, float* s1, float* s2, void* dy, float* dbias, float* dgamma, float* dbeta,
                       float grad_unscale, void* stream);
/* dout, z, dz fp16; dfilm = grad_unscale * (...), dbias += grad_unscale * colsum(dz) */
int vnqa_film_backward_h(const void* dout, const void* z, const float* film, const int32_t* film_row, int32_t film_ld,
                         int32_t alpha_off, int32_t n_img, int32_t C, void* dz, float* dfilm, float* dbias,
                         int32_t mask_relu, float grad_unscale, void* stream);

#ifdef __cplusplus
}
#endif
#endif
