// K1/K2/K3: implicit-GEMM convolution on tcgen05 tensor cores for sm_100a.
//
// Replaces nn.Conv2d 3x3 / 1x1 (+bias, ReLU, BatchNorm affine, FiLM modulation, residual add) of
// models/film_attn_pt_stem.py:211,219,224-241 and the Linear of :244, and their data gradients.
//
//   D[co][pix] = sum_{tap, ci} Wp[tap][co][ci] * X[pix (+) tap][ci]
//
// ONE kernel source for both operand types (template T, optype.cuh):
//   float  - fp32 storage holding tf32-representable values, tcgen05.mma.kind::tf32
//   __half - fp16 storage (the same 10-bit mantissa), tcgen05.mma.kind::f16: twice the MMA rate, half the bytes
// A stage row is one 128-byte swizzle row in both cases (32 fp32 or 64 fp16 reduction elements), one MMA consumes
// 32 bytes of it, so the pipeline - stage bytes, barriers, tile shapes - is identical; fp32 accumulation in TMEM.
//
// Orientation: the UMMA M dimension (TMEM lanes) is the OUTPUT CHANNEL, the N dimension (208 TMEM columns) is a
// block of pixels = 8 frame images x 2 image rows x 13 columns.  130-pixel images do not tile by 128, but
// 13*16 = 208 is a legal UMMA N, so no MMA work is wasted on padding.  The pixel block of filter tap (dy,dx) is a
// 4-D TMA box {32|64 ch, 13 w, 2 h, images} at coordinate {k, dx, h0+dy, n0}; the zero halo of the 3x3 filter is the
// TMA out-of-bounds fill, so activations are stored unpadded.
//
// Two variants of the same pipeline (template NCTA):
//   NCTA = 2 (whenever Cout spans an even number of 128-row tiles): a CTA PAIR on one TPC executes
//     tcgen05.mma.cta_group::2 with M = 256 - each CTA stages its own 128 weight rows and HALF of the pixel block
//     (4 of the 8 images), so per-SM L2->smem traffic drops from 42 KB to 29 KB per K-step and 7 stages fit; only
//     the leader issues MMAs, commits are multicast to both CTAs' barriers, each CTA drains its own TMEM half.
//   NCTA = 1: single-CTA 128 x 208 tiles, 5 stages (odd tile counts: 128-row GEMMs, tail convolutions).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (one thread), warp 2 TMEM allocator, warps 4-7 epilogue
// (TMEM -> registers -> fused math -> coalesced global stores; one thread per output channel).  Accumulators are
// double-buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.  Persistent CTAs.
#include <cstdlib>

#include "common.h"
#include "optype.cuh"
#include "ptx.cuh"

namespace vnqa {

constexpr int BM = 128;            // output channels per CTA tile (TMEM lanes)
constexpr int BN = 208;            // pixels per tile (TMEM columns) = 8 * 2 * 13
constexpr int ROW_BYTES = 128;     // reduction bytes per stage row = one swizzle row (32 fp32 / 64 fp16)
constexpr int A_BYTES = BM * ROW_BYTES;        // 16384
constexpr int ACC_COLS = 256;                  // TMEM column stride between the two accumulator stages
constexpr int TMEM_COLS = 512;
constexpr int IMG_PER_TILE = 8;
constexpr int ROWS_PER_IMG_TILE = 26;          // 2 image rows x 13
constexpr int NUM_THREADS = 256;

template <int NCTA>
struct Cfg {
  static constexpr int B_ROWS = BN / NCTA;                 // pixel rows staged per CTA
  static constexpr int B_BYTES = B_ROWS * ROW_BYTES;       // 26624 / 13312 (multiples of 1024)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 43008 / 29696
  static constexpr int STAGES = NCTA == 1 ? 5 : 7;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers, scheduler ring*/;
};
constexpr int CLC_D = 4;           // depth of the scheduler's response ring (one request is in flight at a time)

struct KParams {
  int taps, k_chunks, m_tiles, n_tiles, splits, rows_pad, flat, cout, n_img, no_round;
  int out_f32;                 // fp16 operands only: BIAS_RELU / PLAIN / BIAS_AFFINE_RELU store fp32 (tensors that leave the fp16 domain)
  // image geometry (conv): W = 13 << g_shift, H = 2 * rp; a tile is (8 >> g_shift) images x 2 rows x W = 208 pixels,
  // i.e. 8 groups of 26 consecutive output rows, (1 << g_shift) groups per image
  int rp, g_shift, hw, w2, pool;
  int dynamic;                 // 1: tiles are handed out by cluster launch control (grid = one cluster per work item)
  long long out_split_stride;  // PARTIAL: n_img * cout
  void* out;
  const float* bias;
  const float* scale;
  const float* shift;
  const int* img_frame;
  double* stat_sum;
  double* stat_sq;
  const float* film;
  const int* film_row;
  int film_ld, alpha_off;
  const void* res;
  void* z_out;
  const void* add;
  float* dbias;
  float grad_unscale;          // DGRAD_RES: dbias += grad_unscale * column sums (1 unless the signals carry a loss scale)
  unsigned long long* dbg;     // VNQA_DEBUG_MBAR builds: host-mapped diagnostics buffer (ptx.cuh), else unused
};

struct TileCoord {
  int mt, nt, sp;
};
// work item -> tile of THIS CTA.  Items enumerate (m-tile group, n-tile, split) with the m index fastest so that
// co-resident CTAs share activation tiles in L2; a CTA pair shares one item and takes m-tiles 2g and 2g+1.
template <int NCTA>
__device__ __forceinline__ TileCoord decode_tile(int item, const KParams& p, int rank) {
  const int mg = p.m_tiles / NCTA;
  TileCoord t;
  t.mt = (item % mg) * NCTA + rank;
  const int r = item / mg;
  t.nt = r % p.n_tiles;
  t.sp = r / p.n_tiles;
  return t;
}

// Tile scheduler.  Every role warp (TMA producer, MMA issuer, epilogue) of every CTA of the cluster walks the same
// sequence of work items.  Static mode: item = cluster id + k * number of clusters (persistent grid).  Dynamic mode
// (cluster launch control): the grid has ONE cluster per work item, the hardware starts as many as fit, and a running
// cluster, instead of exiting after its own item, asks to cancel a not-yet-launched cluster and takes over its item:
// clusters that start late (SMs still busy with the side stream or with NCCL) or run slowly simply take fewer items.
// The leader CTA's producer warp issues ONE request when it starts an item (the response - multicast to both CTAs -
// names the item after it); every role warp consumes each response once and then arrives on the leader's `empty`
// barrier of that ring slot.
template <int NCTA>
struct TileSched {
  static constexpr int CONSUMERS = NCTA == 2 ? 11 : 6;   // leader: producer, MMA, 4 epilogue warps; peer: producer, 4 epilogue
  uint64_t* full;
  uint64_t* empty;
  uint4* resp;
  int dynamic, total, stride;
  int q;        // responses consumed (consumers) / requests issued (the leader's producer keeps a second counter)
  int qi;

  __device__ __forceinline__ int first() const { return (int)blockIdx.x / NCTA; }

  __device__ __forceinline__ void request(int lane) {     // leader producer warp only, warp-uniform
    if (!dynamic) return;
    const int s = qi % CLC_D;
    mbar_wait(&empty[s], (((qi / CLC_D) & 1) ^ 1));
    if (NCTA == 2) { if (lane < 2) mbar_expect_tx_remote(&full[s], 16, lane); }
    else if (lane == 0) mbar_expect_tx(&full[s], 16);
    __syncwarp();
    if (lane == 0) clc_try_cancel<NCTA == 2>(&resp[s], &full[s]);
    ++qi;
  }

  __device__ __forceinline__ int next(int item, int lane, bool leader) {   // warp-uniform
    if (!dynamic) {
      const int n = item + stride;
      return n < total ? n : -1;
    }
    const int s = q % CLC_D;
    mbar_wait(&full[s], (q / CLC_D) & 1);
    const int x = clc_response_ctaid_x(&resp[s]);
    fence_proxy_async();      // this read of the response is ordered before the async-proxy write of a later response
    __syncwarp();
    if (lane == 0) {
      if (leader) mbar_arrive(&empty[s]);
      else mbar_arrive_remote(&empty[s], 0);
    }
    ++q;
    return x < 0 ? -1 : x / NCTA;
  }
};

// ------------------------------------------------------------------------------------------------ epilogue
// One accumulator tile (128 channels x 208 pixels) from TMEM to global memory; thread = output channel.
template <int EPI, typename T>
__device__ __forceinline__ void epilogue_tile(const KParams& p, const TileCoord& tc, uint32_t t_base, int ew,
                                              int lane) {
  using O = Op<T>;
  const int co = tc.mt * BM + ew * 32 + lane;
  const bool co_ok = co < p.cout;
  float bias = 0.f, scale = 1.f, shift = 0.f;
  if (EPI == VNQA_EPI_BIAS_RELU || EPI == VNQA_EPI_BIAS_RELU_AFFINE || EPI == VNQA_EPI_BIAS_RELU_STATS ||
      EPI == VNQA_EPI_FILM || EPI == VNQA_EPI_BIAS || EPI == VNQA_EPI_BIAS_AFFINE_RELU) {
    if (co_ok) bias = __ldg(p.bias + co);
  }
  if ((EPI == VNQA_EPI_BIAS_RELU_AFFINE || EPI == VNQA_EPI_BIAS_AFFINE_RELU) && co_ok) {
    scale = __ldg(p.scale + co);
    shift = __ldg(p.shift + co);
  }
  // activation-typed (T) and fp32 views of the output; which one is used is a property of the epilogue
  T* out_t = reinterpret_cast<T*>(p.out);
  float* out_f = reinterpret_cast<float*>(p.out);
  if (EPI == VNQA_EPI_PARTIAL) out_f += (long long)tc.sp * p.out_split_stride;
  const bool f32_out = O::HALF && p.out_f32;    // fp16 path only; the tf32 path stores fp32 through out_t anyway

  if (EPI == VNQA_EPI_BIAS_AFFINE_RELU && p.pool) {
    // 2x2 max-pool fused behind BatchNorm(eval) + ReLU: a tile holds both rows of a row pair, so every pooling window
    // lies inside one thread's accumulator columns.  Unit = 26 columns of the upper row + the 26 below them.
    const int g = 1 << p.g_shift, half_g = g >> 1;
    const int wo2 = p.w2 >> 2;                       // pooled row length W / 2
#pragma unroll 1
    for (int nl = 0; nl < IMG_PER_TILE; ++nl) {
      if ((nl & (g - 1)) >= half_g) continue;        // lower-row groups are consumed with their upper partner
      uint32_t va[ROWS_PER_IMG_TILE], vb[ROWS_PER_IMG_TILE];
      const uint32_t ca = t_base + nl * ROWS_PER_IMG_TILE, cb = ca + half_g * ROWS_PER_IMG_TILE;
      tmem_ld_x16(ca, va); tmem_ld_x8(ca + 16, va + 16); tmem_ld_x2(ca + 24, va + 24);
      tmem_ld_x16(cb, vb); tmem_ld_x8(cb + 16, vb + 16); tmem_ld_x2(cb + 24, vb + 24);
      tmem_ld_wait();
      const int img = (tc.nt / p.rp) * (IMG_PER_TILE >> p.g_shift) + (nl >> p.g_shift);
      if (img < p.n_img && co_ok) {
        const long long prow0 = (long long)img * (p.hw >> 2) + (tc.nt % p.rp) * wo2 + (nl & (g - 1)) * 13;
        const long long o0 = prow0 * p.cout + co;
#pragma unroll
        for (int j = 0; j < 13; ++j) {
          // relu and max commute; the affine does not (scale may be negative), so it is applied to all four
          const float y0 = fmaf(__uint_as_float(va[2 * j]) + bias, scale, shift);
          const float y1 = fmaf(__uint_as_float(va[2 * j + 1]) + bias, scale, shift);
          const float y2 = fmaf(__uint_as_float(vb[2 * j]) + bias, scale, shift);
          const float y3 = fmaf(__uint_as_float(vb[2 * j + 1]) + bias, scale, shift);
          const float y = fmaxf(fmaxf(fmaxf(y0, y1), fmaxf(y2, y3)), 0.f);
          if (f32_out) out_f[o0 + j * (long long)p.cout] = y;
          else O::st(out_t + o0 + j * (long long)p.cout, y, p.no_round);
        }
      }
      __syncwarp();
    }
    return;
  }

  double st_sum = 0.0, st_sq = 0.0;  // STATS: running per-frame accumulators
  int st_frame = -1;
  float db_acc = 0.f;                // DGRAD_RES: per-channel sum

#pragma unroll 1
  for (int nl = 0; nl < IMG_PER_TILE; ++nl) {
    uint32_t v[ROWS_PER_IMG_TILE];
    tmem_ld_x16(t_base + nl * ROWS_PER_IMG_TILE, v);
    tmem_ld_x8(t_base + nl * ROWS_PER_IMG_TILE + 16, v + 16);
    tmem_ld_x2(t_base + nl * ROWS_PER_IMG_TILE + 24, v + 24);
    tmem_ld_wait();

    long long row0;  // first output row of this 26-column group
    int nvalid;      // valid rows in the group
    int img = 0;
    if (p.flat) {
      row0 = (long long)tc.nt * BN + nl * ROWS_PER_IMG_TILE;
      long long rem = (long long)p.n_img - row0;
      nvalid = rem >= ROWS_PER_IMG_TILE ? ROWS_PER_IMG_TILE : (rem > 0 ? (int)rem : 0);
    } else {
      img = (tc.nt / p.rp) * (IMG_PER_TILE >> p.g_shift) + (nl >> p.g_shift);
      row0 = (long long)img * p.hw + (tc.nt % p.rp) * p.w2 + (nl & ((1 << p.g_shift) - 1)) * ROWS_PER_IMG_TILE;
      nvalid = img < p.n_img ? ROWS_PER_IMG_TILE : 0;
    }
    if (nvalid > 0 && co_ok) {
      const long long o0 = row0 * p.cout + co;
      const long long ld = p.cout;

      if (EPI == VNQA_EPI_BIAS_RELU) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r)
          if (r < nvalid) {
            const float y = fmaxf(__uint_as_float(v[r]) + bias, 0.f);
            if (f32_out) out_f[o0 + r * ld] = y;
            else O::st(out_t + o0 + r * ld, y, p.no_round);
          }
      } else if (EPI == VNQA_EPI_BIAS) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r)
          if (r < nvalid) O::st(out_t + o0 + r * ld, __uint_as_float(v[r]) + bias, p.no_round);
      } else if (EPI == VNQA_EPI_BIAS_AFFINE_RELU) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r)
          if (r < nvalid) {
            const float y = fmaxf(fmaf(__uint_as_float(v[r]) + bias, scale, shift), 0.f);
            if (f32_out) out_f[o0 + r * ld] = y;
            else O::st(out_t + o0 + r * ld, y, p.no_round);
          }
      } else if (EPI == VNQA_EPI_BIAS_RELU_AFFINE) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r)
          if (r < nvalid)
            O::st(out_t + o0 + r * ld, fmaf(fmaxf(__uint_as_float(v[r]) + bias, 0.f), scale, shift), p.no_round);
      } else if (EPI == VNQA_EPI_BIAS_RELU_STATS) {
        const int fr = __ldg(p.img_frame + img);
        if (fr != st_frame) {
          if (st_frame >= 0) {
            atomicAdd(p.stat_sum + (long long)st_frame * p.cout + co, st_sum);
            atomicAdd(p.stat_sq + (long long)st_frame * p.cout + co, st_sq);
          }
          st_frame = fr; st_sum = 0.0; st_sq = 0.0;
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r) {
          const float y = fmaxf(__uint_as_float(v[r]) + bias, 0.f);
          out_f[o0 + r * ld] = y;           // BatchNorm input stays fp32 (normalised by bn_apply)
          s1 += y;
          s2 = fmaf(y, y, s2);
        }
        st_sum += (double)s1;
        st_sq += (double)s2;
      } else if (EPI == VNQA_EPI_FILM) {
        const float* frow = p.film + (long long)__ldg(p.film_row + img) * p.film_ld + p.alpha_off;
        const float alpha = __ldg(frow + co);
        const float beta = __ldg(frow + p.cout + co);
        const T* rrow = reinterpret_cast<const T*>(p.res) + o0;
        float rv[ROWS_PER_IMG_TILE];
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r) rv[r] = O::ld(rrow + r * ld);
        if (p.z_out != nullptr) {
          T* zrow = reinterpret_cast<T*>(p.z_out) + o0;
#pragma unroll
          for (int r = 0; r < ROWS_PER_IMG_TILE; ++r) O::st(zrow + r * ld, __uint_as_float(v[r]) + bias, 1);
        }
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r) {
          const float z = __uint_as_float(v[r]) + bias;
          O::st(out_t + o0 + r * ld, fmaxf(fmaf(alpha, z, beta), 0.f) + rv[r], p.no_round);
        }
      } else if (EPI == VNQA_EPI_PARTIAL) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r)
          if (r < nvalid) out_f[o0 + r * ld] = __uint_as_float(v[r]);
      } else if (EPI == VNQA_EPI_PLAIN) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r)
          if (r < nvalid) {
            if (f32_out) out_f[o0 + r * ld] = __uint_as_float(v[r]);
            else O::st(out_t + o0 + r * ld, __uint_as_float(v[r]), 1);
          }
      } else if (EPI == VNQA_EPI_DGRAD_RES) {
        const T* arow = reinterpret_cast<const T*>(p.add) + o0;
        const T* mrow = reinterpret_cast<const T*>(p.res) + o0;
        float av[ROWS_PER_IMG_TILE], mv[ROWS_PER_IMG_TILE];
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r) {
          av[r] = O::ld(arow + r * ld);
          mv[r] = O::ld(mrow + r * ld);
        }
#pragma unroll
        for (int r = 0; r < ROWS_PER_IMG_TILE; ++r) {
          const float g = mv[r] > 0.f ? O::round(__uint_as_float(v[r]) + av[r]) : 0.f;   // feeds the next MMA
          O::st(out_t + o0 + r * ld, g, 1);
          db_acc += g;
        }
      }
    }  // nvalid && co_ok
    __syncwarp();  // reconverge before the next warp-collective tcgen05.ld
  }
  if (EPI == VNQA_EPI_BIAS_RELU_STATS && st_frame >= 0 && co_ok) {
    atomicAdd(p.stat_sum + (long long)st_frame * p.cout + co, st_sum);
    atomicAdd(p.stat_sq + (long long)st_frame * p.cout + co, st_sq);
  }
  if (EPI == VNQA_EPI_DGRAD_RES && p.dbias != nullptr && co_ok) atomicAdd(p.dbias + co, db_acc * p.grad_unscale);
}

// ------------------------------------------------------------------------------------------------ kernel
template <int EPI, int NCTA, typename T>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
  using C = Cfg<NCTA>;
  using O = Op<T>;
  constexpr int BK = O::BK;   // reduction elements per stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* acc_full = empty_bar + C::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* clc_full = acc_empty + 2;
  uint64_t* clc_empty = clc_full + CLC_D;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(clc_empty + CLC_D);
  uint4* clc_resp = reinterpret_cast<uint4*>(smem + C::STAGES * C::STAGE_BYTES + 256);   // 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int k_iters = p.taps * p.k_chunks;
  TileSched<NCTA> sched{clc_full, clc_empty, clc_resp, p.dynamic, (p.m_tiles / NCTA) * p.n_tiles * p.splits,
                        (int)gridDim.x / NCTA, 0, 0};
  const DbgCtx dbg{p.dbg, full_bar, 2 * C::STAGES + 4};
  (void)dbg;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], NCTA);        // one producer arrival per CTA (+ the TMA byte count)
      mbar_init(&empty_bar[i], 1);          // one (multicast) MMA commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4 * NCTA);   // one arrive per epilogue warp of every CTA
    }
    for (int i = 0; i < CLC_D; ++i) {
      mbar_init(&clc_full[i], 1);           // the expect_tx arrival (+ 16 response bytes)
      mbar_init(&clc_empty[i], TileSched<NCTA>::CONSUMERS);
    }
    fence_mbar_init();
  }
  // A CTA pair allocates tensor memory with a PAIR-wide instruction (UTCATOMSWS.2CTA): both CTAs must be resident
  // and past their barrier initialisation before either issues it.  The CTAs of a cluster can start staggered when
  // the cluster lands on SMs that other work has just vacated (a persistent grid sharing the GPU with another
  // stream); an alloc issued before the peer was up wedged the peer's own alloc for good (round-2 probe,
  // profiles/r02_dp_hang_rootcause.md).  Hence: cluster barrier, THEN alloc, then the usual fence/barrier/fence.
#ifndef VNQA_NO_TMEM_FIX   /* A/B switch for the soak test of the fix (make nofix); never defined in the product */
  if (NCTA == 2) cluster_sync_all();
#endif
  if (warp == 2) {
    if (NCTA == 2) { tmem_alloc_2sm(tmem_slot, TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA)
    int stage = 0;
    uint32_t phase = 0;
    for (int item = sched.first(); item >= 0; item = sched.next(item, lane, leader)) {
      if (leader) sched.request(lane);      // ask now for the item after this one
      const TileCoord tc = decode_tile<NCTA>(item, p, rank);
      const int a_row0 = tc.mt * BM;
      int b1, b2, b3;  // B box coordinates of dims 1..3 before the tap shift
      if (p.flat) {
        b1 = tc.nt * BN + rank * C::B_ROWS; b2 = 0; b3 = 0;
      } else {
        const int ipt = IMG_PER_TILE >> p.g_shift;   // images per tile: 8 / 4 / 2
        b1 = 0; b2 = (tc.nt % p.rp) * 2; b3 = (tc.nt / p.rp) * ipt + rank * (ipt / NCTA);
      }
      const int k_base = tc.sp * p.k_chunks * BK;
      for (int tap = 0; tap < p.taps; ++tap) {
        int dy = 0, dx = 0;
        if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          MBAR_WAIT(&empty_bar[stage], phase ^ 1, 1, ((long long)item << 20) | (tap << 12) | kc);
          if (elect_one()) {
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            if (NCTA == 1) {
              mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
              tma_load_2d(&tmA, &full_bar[stage], sa, k_base + kc * BK, tap * p.rows_pad + a_row0);
              tma_load_4d(&tmB, &full_bar[stage], sb, k_base + kc * BK, b1 + dx, b2 + dy, b3);
            } else {
              // both CTAs' bytes land on the leader's barrier; the peer adds its arrival remotely
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
              tma_load_2d_2sm(&tmA, &full_bar[stage], sa, k_base + kc * BK, tap * p.rows_pad + a_row0);
              tma_load_4d_2sm(&tmB, &full_bar[stage], sb, k_base + kc * BK, b1 + dx, b2 + dy, b3);
              if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    constexpr uint32_t idesc = O::idesc(BM * NCTA, BN, 0, 0);
    constexpr uint64_t dbase = umma_desc_base_sw128(16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = sched.first(); item >= 0; item = sched.next(item, lane, true)) {
      MBAR_WAIT(&acc_empty[as], aphase ^ 1, 2, ((long long)item << 20));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * ACC_COLS;
      for (int it = 0; it < k_iters; ++it) {
        MBAR_WAIT(&full_bar[stage], phase, 3, ((long long)item << 20) | it);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = umma_desc(dbase, sa);
          const uint64_t bdesc = umma_desc(dbase, sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < ROW_BYTES / 32; ++k) {
            // advance 32 bytes (>>4 = 2) inside the 128-byte swizzle row
            O::template mma<NCTA>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          if (NCTA == 1) {
            umma_commit(&empty_bar[stage]);                       // frees the smem stage when the MMAs retire
            if (it == k_iters - 1) umma_commit(&acc_full[as]);    // accumulator complete -> epilogue
          } else {
            umma_commit_2sm(&empty_bar[stage], 0x3);              // ... in both CTAs of the pair
            if (it == k_iters - 1) umma_commit_2sm(&acc_full[as], 0x3);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (every CTA, own TMEM half)
    const int ew = warp - 4;  // TMEM lanes [32*ew, 32*ew+32)
    int as = 0;
    uint32_t aphase = 0;
    for (int item = sched.first(); item >= 0; item = sched.next(item, lane, leader)) {
      const TileCoord tc = decode_tile<NCTA>(item, p, rank);
      MBAR_WAIT(&acc_full[as], aphase, 4, ((long long)item << 20));
      tc_fence_after();
      const uint32_t t_base = tmem_base + (uint32_t(ew * 32) << 16) + as * ACC_COLS;
      epilogue_tile<EPI, T>(p, tc, t_base, ew, lane);
      // all TMEM reads of this accumulator stage are done -> hand it back to the (leader's) MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&acc_empty[as]);
        else mbar_arrive_remote(&acc_empty[as], 0);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();   // the leader's MMAs read the peer's smem until the end
  if (warp == 2) {
    if (NCTA == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  // neither CTA of a pair leaves (and frees its SM for the next cluster) before BOTH have released their tensor memory
#ifndef VNQA_NO_TMEM_FIX
  if (NCTA == 2) cluster_sync_all();
#endif
}

template <int EPI, int NCTA, typename T>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& kp, int sms, cudaStream_t st) {
  auto kern = conv_gemm_kernel<EPI, NCTA, T>;
  constexpr int smem = Cfg<NCTA>::SMEM_BYTES;
  static bool attr_set[64] = {false};      // per instantiation and device: the attribute is sticky
  int dev = 0;
  VNQA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    VNQA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int items = (kp.m_tiles / NCTA) * kp.n_tiles * kp.splits;
  const int groups = sms / NCTA;
  // dynamic: one cluster per work item, handed out by cluster launch control; static: persistent grid, round robin
  const int grid = (kp.dynamic || items < groups ? items : groups) * NCTA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = NCTA == 2 ? 1 : 0;
  VNQA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, kp));
  return 0;
}

template <int EPI, typename T>
static int launch_any(int ncta, const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& kp, int sms,
                      cudaStream_t st) {
  return ncta == 2 ? launch<EPI, 2, T>(tmA, tmB, kp, sms, st) : launch<EPI, 1, T>(tmA, tmB, kp, sms, st);
}

// A/B switch VNQA_TILE_SCHED=static (persistent grid with the round-robin schedule of round 1), read once
static int dynamic_sched() {
  static const int v = [] { const char* e = getenv("VNQA_TILE_SCHED"); return (e && e[0] == 's') ? 0 : 1; }();
  return v;
}

// bring-up switch VNQA_CONV_NCTA=1 (single-CTA kernels everywhere), read once
static int forced_ncta() {
  static const int v = [] { const char* e = getenv("VNQA_CONV_NCTA"); return e ? atoi(e) : 0; }();
  return v;
}

template <typename T>
static int conv_gemm_typed(const vnqa_conv_args* a, cudaStream_t st) {
  using O = Op<T>;
  constexpr int BK = O::BK;
  constexpr int ES = O::BYTES;
  // fp16: a reduction that is not a multiple of 64 is completed by the TMA out-of-bounds zero fill (both operands)
  VNQA_REQUIRE(a->cin > 0 && (a->cin * ES) % 16 == 0, "vnqa_conv_gemm: cin must be a positive multiple of %d (got %d)",
               16 / ES, a->cin);
  const int splits = a->flat ? (a->splits > 0 ? a->splits : 1) : 1;
  VNQA_REQUIRE(splits == 1 || a->cin % BK == 0, "vnqa_conv_gemm: split-K needs cin per split to be a multiple of %d", BK);
  VNQA_REQUIRE(O::HALF || a->cin % BK == 0, "vnqa_conv_gemm: cin must be a multiple of %d (got %d)", BK, a->cin);
  VNQA_REQUIRE(!a->out_f32 || (O::HALF && (a->epi == VNQA_EPI_BIAS_RELU || a->epi == VNQA_EPI_PLAIN ||
                                           a->epi == VNQA_EPI_BIAS_AFFINE_RELU)),
               "vnqa_conv_gemm: out_f32 is available for fp16 operands with BIAS_RELU, PLAIN and BIAS_AFFINE_RELU");

  // image geometry: 10 x 13 (the FiLM trunk) unless the caller names one of the frozen stem's (models/obj_detector.py:
  // 40 x 52 and 20 x 26): W = 13 * 2^s keeps the tile at 208 pixels = (8 >> s) images x 2 rows x W
  const int img_h = a->img_h > 0 ? a->img_h : VNQA_IMG_H;
  const int img_w = a->img_w > 0 ? a->img_w : VNQA_IMG_W;
  int g_shift = 0;
  if (!a->flat) {
    VNQA_REQUIRE(img_w == 13 || img_w == 26 || img_w == 52, "vnqa_conv_gemm: img_w must be 13, 26 or 52 (got %d)", img_w);
    VNQA_REQUIRE(img_h >= 2 && img_h % 2 == 0, "vnqa_conv_gemm: img_h must be even (got %d)", img_h);
    g_shift = img_w == 13 ? 0 : (img_w == 26 ? 1 : 2);
    VNQA_REQUIRE(!a->pool || (a->epi == VNQA_EPI_BIAS_AFFINE_RELU && g_shift > 0),
                 "vnqa_conv_gemm: pool needs BIAS_AFFINE_RELU and img_w 26 or 52");
  } else {
    VNQA_REQUIRE(!a->pool, "vnqa_conv_gemm: pool on a flat GEMM");
  }
  const int ipt = IMG_PER_TILE >> g_shift;

  KParams kp{};
  kp.taps = a->taps;
  kp.k_chunks = (a->cin + BK - 1) / BK;
  kp.m_tiles = a->rows_pad / BM;
  kp.rp = img_h / 2; kp.g_shift = g_shift; kp.hw = img_h * img_w; kp.w2 = 2 * img_w; kp.pool = a->pool;
  kp.n_tiles = a->flat ? (a->n_img + BN - 1) / BN : ((a->n_img + ipt - 1) / ipt) * kp.rp;
  kp.splits = splits;
  kp.rows_pad = a->rows_pad;
  kp.flat = a->flat;
  kp.cout = a->cout;
  kp.n_img = a->n_img;
  kp.no_round = a->no_round;
  kp.out_f32 = a->out_f32;
  kp.out_split_stride = (long long)a->n_img * a->cout;
  kp.out = a->out; kp.bias = a->bias; kp.scale = a->scale; kp.shift = a->shift;
  kp.img_frame = a->img_frame; kp.stat_sum = a->stat_sum; kp.stat_sq = a->stat_sq;
  kp.film = a->film; kp.film_row = a->film_row; kp.film_ld = a->film_ld; kp.alpha_off = a->alpha_off;
  kp.res = a->res; kp.z_out = a->z_out; kp.add = a->add; kp.dbias = a->dbias;
  kp.grad_unscale = a->grad_unscale == 0.f ? 1.f : a->grad_unscale;
  kp.dbg = debug_buffer();
  kp.dynamic = dynamic_sched();

  // CTA pairs (cta_group::2) whenever the output channels split into an even number of 128-row tiles
  int ncta = (kp.m_tiles % 2 == 0) ? 2 : 1;
  if (forced_ncta() == 1) ncta = 1;

  // A: packed weights, 2-D {K, rows}
  CUtensorMap tmA, tmB;
  {
    const uint64_t ktot = a->flat ? (uint64_t)a->cin * splits : (uint64_t)a->cin;
    const uint64_t dims[2] = {ktot, (uint64_t)a->taps * a->rows_pad};
    const uint64_t strides[1] = {ktot * ES};
    const uint32_t box[2] = {BK, BM};
    if (make_tmap(&tmA, a->w_packed, O::HALF, 2, dims, strides, box)) return -3;
  }
  if (a->flat) {
    const uint64_t ld = a->x_ld > 0 ? (uint64_t)a->x_ld : (uint64_t)a->cin * splits;
    VNQA_REQUIRE(ld >= (uint64_t)a->cin * splits && (ld * ES) % 16 == 0, "vnqa_conv_gemm: bad x_ld");
    const uint64_t dims[4] = {(uint64_t)a->cin * splits, (uint64_t)a->n_img, 1, 1};
    const uint64_t strides[3] = {ld * ES, ld * ES * (uint64_t)a->n_img, ld * ES * (uint64_t)a->n_img};
    const uint32_t box[4] = {BK, (uint32_t)(BN / ncta), 1, 1};
    if (make_tmap(&tmB, a->x, O::HALF, 4, dims, strides, box)) return -3;
  } else {
    const uint64_t c = (uint64_t)a->cin;
    const uint64_t dims[4] = {c, (uint64_t)img_w, (uint64_t)img_h, (uint64_t)a->n_img};
    const uint64_t strides[3] = {c * ES, c * ES * img_w, c * ES * img_w * img_h};
    const uint32_t box[4] = {BK, (uint32_t)img_w, 2, (uint32_t)(ipt / ncta)};
    if (make_tmap(&tmB, a->x, O::HALF, 4, dims, strides, box)) return -3;
  }

  int sms = vnqa_sm_count();
  if (sms <= 0) return -2;
  switch (a->epi) {
    case VNQA_EPI_BIAS_RELU:
      VNQA_REQUIRE(a->bias, "conv_gemm: BIAS_RELU needs bias");
      return launch_any<VNQA_EPI_BIAS_RELU, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_BIAS_RELU_AFFINE:
      VNQA_REQUIRE(a->bias && a->scale && a->shift, "conv_gemm: AFFINE needs bias/scale/shift");
      return launch_any<VNQA_EPI_BIAS_RELU_AFFINE, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_BIAS_RELU_STATS:
      VNQA_REQUIRE(!a->flat && a->bias && a->img_frame && a->stat_sum && a->stat_sq, "conv_gemm: STATS operands");
      return launch_any<VNQA_EPI_BIAS_RELU_STATS, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_FILM:
      VNQA_REQUIRE(!a->flat && a->bias && a->film && a->film_row && a->res, "conv_gemm: FILM operands");
      return launch_any<VNQA_EPI_FILM, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_PARTIAL:
      return launch_any<VNQA_EPI_PARTIAL, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_PLAIN:
      VNQA_REQUIRE(splits == 1, "conv_gemm: PLAIN cannot split K");
      return launch_any<VNQA_EPI_PLAIN, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_DGRAD_RES:
      VNQA_REQUIRE(!a->flat && a->add && a->res, "conv_gemm: DGRAD_RES operands");
      return launch_any<VNQA_EPI_DGRAD_RES, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_BIAS:
      VNQA_REQUIRE(a->bias, "conv_gemm: BIAS needs bias");
      return launch_any<VNQA_EPI_BIAS, T>(ncta, tmA, tmB, kp, sms, st);
    case VNQA_EPI_BIAS_AFFINE_RELU:
      VNQA_REQUIRE(!a->flat && a->bias && a->scale && a->shift, "conv_gemm: BIAS_AFFINE_RELU needs bias/scale/shift");
      return launch_any<VNQA_EPI_BIAS_AFFINE_RELU, T>(ncta, tmA, tmB, kp, sms, st);
    default:
      set_error("vnqa_conv_gemm: unknown epilogue %d", a->epi);
      return -1;
  }
}

}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_conv_gemm(const vnqa_conv_args* a, void* stream) {
  VNQA_REQUIRE(a != nullptr, "vnqa_conv_gemm: null args");
  VNQA_REQUIRE(a->taps == 1 || a->taps == 9, "vnqa_conv_gemm: taps must be 1 or 9 (got %d)", a->taps);
  VNQA_REQUIRE(a->rows_pad % BM == 0 && a->rows_pad >= a->cout, "vnqa_conv_gemm: rows_pad must be cout rounded up to %d", BM);
  VNQA_REQUIRE(a->n_img > 0, "vnqa_conv_gemm: empty batch");
  VNQA_REQUIRE(a->w_packed && a->x && a->out, "vnqa_conv_gemm: null operand");
  VNQA_REQUIRE(!(a->flat && a->taps != 1), "vnqa_conv_gemm: flat GEMM has one tap");
  VNQA_REQUIRE(a->act_dtype == VNQA_ACT_TF32 || a->act_dtype == VNQA_ACT_FP16, "vnqa_conv_gemm: unknown act_dtype %d",
               a->act_dtype);
  return a->act_dtype == VNQA_ACT_FP16 ? conv_gemm_typed<__half>(a, as_stream(stream))
                                       : conv_gemm_typed<float>(a, as_stream(stream));
}
