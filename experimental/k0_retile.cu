// EXPERIMENTAL (not built, not validated on hardware yet): K0 re-tiled for DRAM locality.
// v[b][c][p][t] (t innermost) -> x[img(b,t)][p][c].  For a fixed (b, c) the 13 pixels of one image row x T frames are
// 13*T contiguous floats (1.8 KB at T = 35): a block stages [32 c][13 p][T t] through shared memory (58 KB) with long
// contiguous reads and writes 128-byte channel runs.  Today's kernel (layout.cu) reads 140-byte runs and reaches
// 0.35 of the HBM roofline.
#include <cstdint>
#include <cuda_runtime.h>

namespace vnqa_exp {

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// grid (10 image rows, C/32 slabs, B clips), 256 threads, dynamic smem 32*13*T floats (+ padding)
__global__ void __launch_bounds__(256) frames_to_nhwc_rows_kernel(const float* __restrict__ v, float* __restrict__ x,
                                                                  const int* __restrict__ img_of, int C, int T) {
  extern __shared__ float tile[];             // [32][13*T + 1]
  const int row = blockIdx.x, cs = blockIdx.y, b = blockIdx.z;
  const int run = 13 * T;                     // contiguous floats per channel
  const int ld = run + 1;
  const float* src = v + (((long long)b * C + cs * 32) * 130 + row * 13) * T;
  for (int i = threadIdx.x; i < 32 * run; i += blockDim.x) {
    const int c = i / run, k = i - c * run;
    tile[c * ld + k] = __ldg(src + (long long)c * 130 * T + k);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int pt = threadIdx.x >> 5; pt < run; pt += blockDim.x >> 5) {   // pt = p_local * T + t
    const int pl = pt / T, t = pt - pl * T;
    const int img = __ldg(img_of + b * T + t);
    if (img >= 0) x[((long long)img * 130 + row * 13 + pl) * C + cs * 32 + lane] = tf32_round(tile[lane * ld + pt]);
  }
}

}  // namespace vnqa_exp
