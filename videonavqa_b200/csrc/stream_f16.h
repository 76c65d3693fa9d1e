// fp16-activation variants of the streaming kernels (stream_f16.cu), called by the act_dtype dispatch in layout.cu /
// small.cu.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace vnqa {
int pack_conv_weight_h(const float* w, void* packed, int cout, int cin, int taps, int rows_pad, int transpose_flip,
                       cudaStream_t st);
int bn_apply_h(const float* y, void* x, const int32_t* img_frame, const float* scale, const float* shift, int n_img,
               int C, cudaStream_t st);
int bn_backward_h(const void* dx, const float* y, const int32_t* img_frame, const int32_t* frame_ct, int n_frames,
                  const float* mean, const float* rstd, const float* gamma, int n_img, int C, float* s1, float* s2,
                  void* dy, float* dbias, float* dgamma, float* dbeta, float grad_unscale, cudaStream_t st);
int film_backward_h(const void* dout, const void* z, const float* film, const int32_t* film_row, int film_ld,
                    int alpha_off, int n_img, int C, void* dz, float* dfilm, float* dbias, int mask_relu,
                    float grad_unscale, cudaStream_t st);
}  // namespace vnqa
