// K4: weight gradients on tcgen05 tensor cores (autograd of the convolutions / Linear of
// models/film_attn_pt_stem.py:211,219,224,244).
//
//   dW[tap][co][ci] = grad_unscale * sum_{pixel rows r} dY[r][co] * X[r (+) tap][ci]
//
// The reduction runs over PIXELS, and both operands are stored pixel-major ([row][channel], channel contiguous),
// i.e. "MN-major" for the MMA: no transposed copies are made - the UMMA descriptors are built MN-major
// (instruction-descriptor bits 15/16) over TMA boxes of whole image rows.  ONE source for both operand types:
//   float  (tf32): boxes {32 ch x 104 pixel rows} (104 = 4 images x 2 rows x 13 = 13 K-steps of 8), smem layout
//                  "128B swizzle / 32B atoms" - the only one tf32 accepts MN-major;
//   __half (fp16): boxes {64 ch x 208 pixel rows} (208 = 8 images x 2 rows x 13 = 13 K-steps of 16), ordinary 128B
//                  swizzle (for 16-bit types it IS the canonical MN-major layout).
// Stage bytes are the same (2 stages of 104 KB).  Tap shifts and the zero halo are again TMA box coordinates /
// out-of-bounds fill.  fp32 accumulators in TMEM (double-buffered).
//
// Two variants (template NCTA):
//   NCTA = 2 (Cout and Cin multiples of 256): a CTA pair executes tcgen05.mma.cta_group::2 on a 256 (co) x 256 (ci)
//     tile - each CTA stages its own 128 output channels of dY and its own 128 input channels of X, halving the
//     operand bytes per MMA flop with respect to the single-CTA 128 x 128 tile;
//   NCTA = 1: 128 x 128 tiles (small / odd channel counts, the Linear of :244).
// Split-K over pixel chunks: as many splits as fill the SMs (about two rounds of tiles per CTA group) and at least
// enough to keep one accumulation chain <= 29120 pixel rows (the tensor core's fp32 accumulation truncates, so its
// bias grows with the chain length), at most 16.  SPLIT-MAJOR tile order: the co-resident CTAs always stream the same
// pixel range, so dY / X are read from DRAM about once per round instead of once per tile.
//
// DETERMINISTIC reduction of the split-K partial tiles: every (tile, CTA, epilogue warp) owns a turn counter; split s
// of a tile waits until the counter reads s, then accumulates its partial into the [tap][Cout][Cin] buffer with plain
// loads/stores (split 0 stores, so the buffer needs no zero-fill) and passes the turn on.  Items are issued in
// split-major order, so a predecessor is always an earlier item of the static schedule (owned by a CTA that is
// resident - the grid is persistent and never exceeds the SM count): the wait cannot dead-lock, and with <= 16 splits
// the predecessor's epilogue has usually retired by the time the successor's main loop ends.  The order of the fp32
// additions is therefore fixed: two launches on the same operands give the same bits.  (Round 1 used red.global.add.)
// The partial tile is transposed through shared memory so that every warp access hits 128 contiguous bytes; a final
// pass writes PyTorch's (Cout, Cin, kh, kw).
#include <cstdlib>

#include "common.h"
#include "optype.cuh"
#include "ptx.cuh"

namespace vnqa {

constexpr int WG_M = 128;                        // output channels per CTA (TMEM lanes)
constexpr int WG_NB = 128;                       // input channels STAGED per CTA
constexpr int WG_STAGES = 2;
constexpr int WG_SCRATCH = 4 * 32 * 33 * 4;      // per-epilogue-warp 32x33 transpose tile
constexpr int WG_THREADS = 256;

template <typename T>
struct WgCfg {
  static constexpr int CH = Op<T>::BK;                       // channels per TMA box (128 bytes): 32 / 64
  static constexpr int IMG = Op<T>::HALF ? 8 : 4;            // images per pixel chunk
  static constexpr int KP = IMG * 26;                        // pixel rows per stage: 104 / 208
  static constexpr int BOX_BYTES = KP * 128;                 // 13312 / 26624
  static constexpr int A_BYTES = (WG_M / CH) * BOX_BYTES;    // 53248
  static constexpr int B_BYTES = (WG_NB / CH) * BOX_BYTES;   // 53248
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 106496
  static constexpr int KSTEP_ROWS = Op<T>::UMMA_K;           // pixel rows per MMA: 8 / 16
  static constexpr int KSTEP_DESC = KSTEP_ROWS * 128 / 16;   // descriptor start-address advance per K-step: 64 / 128
  static constexpr int SMEM = WG_STAGES * STAGE_BYTES + WG_SCRATCH + 1024 + 256;
  __host__ __device__ static constexpr uint64_t dbase() {
    // LBO = next box of channels; SBO = next group of K rows (4 rows of 128 B for 32-byte atoms, 8 rows otherwise)
    return Op<T>::HALF ? umma_desc_base_mn_sw128(BOX_BYTES, 1024) : umma_desc_base_sw128_32b(BOX_BYTES, 512);
  }
};

struct WParams {
  int taps, m_tiles, n_tiles, splits, n_chunks, chunks_per_split, flat;
  int cout, cin;          // real sizes (tiles beyond are skipped on store)
  float* out;             // [taps][cout][cin] accumulation buffer (== dw when taps == 1)
  int* turn;              // [base_items][NCTA][4] turn counters, zero on entry, zero again on exit
  float grad_unscale;
  unsigned long long* dbg;   // VNQA_DEBUG_MBAR builds only
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// m_tiles counts 128-row co tiles, n_tiles counts (128*NCTA)-wide ci tiles.
template <int NCTA, typename T>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WParams p) {
  using W = WgCfg<T>;
  using O = Op<T>;
  constexpr int N_INSTR = WG_NB * NCTA;          // MMA N = accumulator columns per CTA
  constexpr int TMEM_COLS = 2 * N_INSTR;         // two accumulator stages (256 / 512 columns)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* scratch = reinterpret_cast<float*>(smem + WG_STAGES * W::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * W::STAGE_BYTES + WG_SCRATCH);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* acc_full = empty_bar + WG_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int mg = p.m_tiles / NCTA;               // co tile groups (a pair takes co tiles 2g, 2g+1)
  const int base_items = mg * p.n_tiles * p.taps;
  const int total_items = base_items * p.splits;
  const int first_item = blockIdx.x / NCTA, item_stride = gridDim.x / NCTA;
  const DbgCtx dbg{p.dbg, full_bar, 2 * WG_STAGES + 4};
  (void)dbg;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full_bar[i], NCTA); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * NCTA); }
    fence_mbar_init();
  }
  // pair-wide TMEM allocation only after both CTAs of the pair are up (see conv_gemm.cu)
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

  // item -> (split, tap, nt, m group), split outermost / m fastest: one wave of CTAs shares one pixel range, CTAs with
  // the same nt share the X slab, CTAs with the same co tile share the dY slab (all through L2)
  auto decode = [&](int item, int& mt, int& nt, int& tap, int& sp) {
    sp = item / base_items; item -= sp * base_items;
    mt = (item % mg) * NCTA + rank; item /= mg;
    nt = item % p.n_tiles;
    tap = item / p.n_tiles;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA)
    int stage = 0; uint32_t phase = 0;
    for (int item = first_item; item < total_items; item += item_stride) {
      int mt, nt, tap, sp;
      decode(item, mt, nt, tap, sp);
      int dy = 0, dx = 0;
      if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
      const int c_begin = sp * p.chunks_per_split;
      const int c_end = min(c_begin + p.chunks_per_split, p.n_chunks);
      const int ci0 = nt * N_INSTR + rank * WG_NB;   // this CTA's 128 input channels of the tile
      for (int c = c_begin; c < c_end; ++c) {
        MBAR_WAIT(&empty_bar[stage], phase ^ 1, 1, ((long long)item << 20) | c);
        if (elect_one()) {
          uint8_t* sa = smem + stage * W::STAGE_BYTES;
          uint8_t* sb = sa + W::A_BYTES;
          int a1, a2, a3, b1, b2, b3;
          if (p.flat) { a1 = c * W::KP; a2 = 0; a3 = 0; b1 = a1; b2 = 0; b3 = 0; }
          else { a1 = 0; a2 = (c % 5) * 2; a3 = (c / 5) * W::IMG; b1 = dx; b2 = a2 + dy; b3 = a3; }
          if (NCTA == 1) {
            mbar_expect_tx(&full_bar[stage], W::STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < WG_M / W::CH; ++i)
              tma_load_4d(&tmA, &full_bar[stage], sa + i * W::BOX_BYTES, mt * WG_M + i * W::CH, a1, a2, a3);
#pragma unroll
            for (int i = 0; i < WG_NB / W::CH; ++i)
              tma_load_4d(&tmB, &full_bar[stage], sb + i * W::BOX_BYTES, ci0 + i * W::CH, b1, b2, b3);
          } else {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * W::STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < WG_M / W::CH; ++i)
              tma_load_4d_2sm(&tmA, &full_bar[stage], sa + i * W::BOX_BYTES, mt * WG_M + i * W::CH, a1, a2, a3);
#pragma unroll
            for (int i = 0; i < WG_NB / W::CH; ++i)
              tma_load_4d_2sm(&tmB, &full_bar[stage], sb + i * W::BOX_BYTES, ci0 + i * W::CH, b1, b2, b3);
            if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
          }
        }
        __syncwarp();
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    constexpr uint32_t idesc = O::idesc(WG_M * NCTA, N_INSTR, 1, 1);       // both operands MN-major
    constexpr uint64_t dbase = W::dbase();
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    for (int item = first_item; item < total_items; item += item_stride) {
      int mt, nt, tap, sp;
      decode(item, mt, nt, tap, sp);
      const int c_begin = sp * p.chunks_per_split;
      const int c_end = min(c_begin + p.chunks_per_split, p.n_chunks);
      MBAR_WAIT(&acc_empty[as], aphase ^ 1, 2, ((long long)item << 20));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * N_INSTR;
      for (int c = c_begin; c < c_end; ++c) {
        MBAR_WAIT(&full_bar[stage], phase, 3, ((long long)item << 20) | c);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * W::STAGE_BYTES);
          const uint64_t adesc = umma_desc(dbase, sa);
          const uint64_t bdesc = umma_desc(dbase, sa + W::A_BYTES);
#pragma unroll
          for (int g = 0; g < W::KP / W::KSTEP_ROWS; ++g)
            O::template mma<NCTA>(d_tmem, adesc + W::KSTEP_DESC * g, bdesc + W::KSTEP_DESC * g, idesc,
                                  (c > c_begin || g > 0) ? 1u : 0u);
          if (NCTA == 1) {
            umma_commit(&empty_bar[stage]);
            if (c == c_end - 1) umma_commit(&acc_full[as]);
          } else {
            umma_commit_2sm(&empty_bar[stage], 0x3);
            if (c == c_end - 1) umma_commit_2sm(&acc_full[as], 0x3);
          }
        }
        __syncwarp();
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (every CTA, own 128 co)
    const int ew = warp - 4;
    int as = 0; uint32_t aphase = 0;
    for (int item = first_item; item < total_items; item += item_stride) {
      int mt, nt, tap, sp;
      decode(item, mt, nt, tap, sp);
      MBAR_WAIT(&acc_full[as], aphase, 4, ((long long)item << 20));
      tc_fence_after();
      // my turn?  (split sp of this tile accumulates after split sp-1 - fixed fp32 addition order)
      int* turn = p.turn + (((item - sp * base_items) * NCTA + rank) * 4 + ew);
      if (sp > 0) {
        if (lane == 0)
          while (ld_acquire(turn) != sp) __nanosleep(40);
        __syncwarp();
      }
      const uint32_t t_base = tmem_base + (uint32_t(ew * 32) << 16) + as * N_INSTR;
      float* tile_s = scratch + ew * (32 * 33);
      const int co0 = mt * WG_M + ew * 32;           // this warp owns output channels co0 .. co0+31
#pragma unroll 1
      for (int j0 = 0; j0 < N_INSTR; j0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(t_base + j0, v);
        tmem_ld_wait();
        // lane = co holds 32 consecutive ci: transpose through smem so that a warp touches 32 consecutive ci of ONE co
#pragma unroll
        for (int j = 0; j < 32; ++j) tile_s[lane * 33 + j] = __uint_as_float(v[j]) * p.grad_unscale;
        __syncwarp();
        const int ci = nt * N_INSTR + j0 + lane;
        if (ci < p.cin) {
          float* obase = p.out + ((long long)tap * p.cout + co0) * p.cin + ci;
          float acc[32];
          if (sp == 0) {
#pragma unroll
            for (int r = 0; r < 32; ++r) acc[r] = 0.f;
          } else {
            // 32 independent L2 loads in flight (the previous split ran on another SM: bypass L1)
#pragma unroll
            for (int r = 0; r < 32; ++r) acc[r] = co0 + r < p.cout ? __ldcg(obase + (long long)r * p.cin) : 0.f;
          }
#pragma unroll
          for (int r = 0; r < 32; ++r)
            if (co0 + r < p.cout) obase[(long long)r * p.cin] = acc[r] + tile_s[r * 33 + lane];
        }
        __syncwarp();
      }
      // publish the tile, pass the turn on (the last split leaves the counter at zero for the next launch)
      __threadfence();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        st_release(turn, sp + 1 == p.splits ? 0 : sp + 1);
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

// [tap][co][ci] -> (co, ci, tap)
__global__ void wgrad_unpack_kernel(const float* __restrict__ ws, float* __restrict__ dw, int cout, int cin, int taps) {
  const long long total = (long long)cout * cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const long long r = i / taps;   // co * cin + ci
    dw[i] = __ldg(ws + (long long)tap * cout * cin + r);
  }
}

template <int NCTA, typename T>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const WParams& wp, int sms, cudaStream_t st) {
  auto kern = wgrad_kernel<NCTA, T>;
  constexpr int smem = WgCfg<T>::SMEM;
  static bool attr_set[64] = {false};
  int dev = 0;
  VNQA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    VNQA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int items = (wp.m_tiles / NCTA) * wp.n_tiles * wp.taps * wp.splits;
  const int groups = sms / NCTA;
  const int grid = (items < groups ? items : groups) * NCTA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = NCTA == 2 ? 1 : 0;
  VNQA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, wp));
  return 0;
}

static int wgrad_forced_ncta() {
  static const int v = [] { const char* e = getenv("VNQA_WGRAD_NCTA"); return e ? atoi(e) : 0; }();
  return v;
}

struct WgShape {
  int ncta, m_tiles, n_tiles, base_items;
  long long counter_bytes;    // leading bytes of ws that must be zero on entry (256-byte multiple)
};
static WgShape wg_shape(int cout, int cin, int taps) {
  WgShape s;
  s.ncta = (cout % 256 == 0 && cin % 256 == 0) ? 2 : 1;
  if (wgrad_forced_ncta() == 1) s.ncta = 1;
  s.m_tiles = (cout + WG_M - 1) / WG_M;
  s.n_tiles = (cin + WG_NB * s.ncta - 1) / (WG_NB * s.ncta);
  s.base_items = (s.m_tiles / s.ncta) * s.n_tiles * taps;
  s.counter_bytes = ((long long)s.base_items * s.ncta * 4 * 4 + 255) / 256 * 256;
  return s;
}

template <typename T>
static int conv_wgrad_typed(const void* dy, const void* x, float* dw, void* ws, int n_img, int cout, int cin, int taps,
                            int flat, float grad_unscale, cudaStream_t st) {
  using W = WgCfg<T>;
  constexpr int ES = Op<T>::BYTES;
  // fp16: channel counts that are not multiples of 64 are completed by the TMA out-of-bounds zero fill
  const int ch_mult = Op<T>::HALF ? 8 : 32;
  VNQA_REQUIRE(cout % ch_mult == 0 && cin % ch_mult == 0, "vnqa_conv_wgrad: channels must be multiples of %d (%d, %d)",
               ch_mult, cout, cin);
  const WgShape sh = wg_shape(cout, cin, taps);
  const int ncta = sh.ncta;
  WParams wp{};
  wp.taps = taps;
  wp.m_tiles = sh.m_tiles;
  wp.n_tiles = sh.n_tiles;
  wp.flat = flat;
  wp.cout = cout;
  wp.cin = cin;
  wp.n_chunks = flat ? (n_img + W::KP - 1) / W::KP : ((n_img + W::IMG - 1) / W::IMG) * 5;
  wp.turn = reinterpret_cast<int*>(ws);
  wp.out = taps == 1 ? dw : reinterpret_cast<float*>(static_cast<char*>(ws) + sh.counter_bytes);
  wp.grad_unscale = grad_unscale == 0.f ? 1.f : grad_unscale;
  wp.dbg = debug_buffer();
  const int sms = vnqa_sm_count();
  if (sms <= 0) return -2;
  // split-K over pixel chunks.  Two lower bounds on the split count:
  //  (a) fill: about two rounds of tiles per CTA group (items are issued split-major, so the ~groups concurrently
  //      running items always stream the same pixel range through L2 - dY / X are read from DRAM about once per round);
  //  (b) precision: the tensor core accumulates in fp32 with truncation, a biased error that grows linearly with the
  //      length of one accumulation chain (measured: 1.2e-5 normalised at 5.8k pixel rows, 6.9e-4 at 291k); chains are
  //      therefore bounded to ~29k pixel rows (<= 1e-4) and the partial tiles are added in round-to-nearest fp32.
  // The turn-ordered reduction serialises the epilogues of one tile's splits, so no more splits than that (<= 16).
  // VNQA_WGRAD_SPLITS overrides (bring-up).
  const int groups = sms / ncta;
  (void)ES;
  constexpr int MAX_CHAIN_CHUNKS = 29120 / W::KP;       // 280 (tf32) / 140 (fp16) chunks
  int splits = (2 * groups) / sh.base_items;
  const int by_chain = (wp.n_chunks + MAX_CHAIN_CHUNKS - 1) / MAX_CHAIN_CHUNKS;
  if (splits < by_chain) splits = by_chain;
  if (splits > 16) splits = 16;
  static const int forced_splits = [] { const char* e = getenv("VNQA_WGRAD_SPLITS"); return e ? atoi(e) : 0; }();
  if (forced_splits > 0) splits = forced_splits;
  if (splits < 1) splits = 1;
  if (splits > wp.n_chunks) splits = wp.n_chunks;
  const int cps = (wp.n_chunks + splits - 1) / splits;
  wp.chunks_per_split = cps;
  wp.splits = (wp.n_chunks + cps - 1) / cps;

  CUtensorMap tmA, tmB;
  auto make = [&](CUtensorMap* m, const void* base, int C) -> int {
    if (flat) {
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)n_img, 1, 1};
      const uint64_t strides[3] = {(uint64_t)C * ES, (uint64_t)C * ES * n_img, (uint64_t)C * ES * n_img};
      const uint32_t box[4] = {(uint32_t)W::CH, (uint32_t)W::KP, 1, 1};
      return make_tmap(m, base, Op<T>::HALF, 4, dims, strides, box, true);
    }
    const uint64_t dims[4] = {(uint64_t)C, VNQA_IMG_W, VNQA_IMG_H, (uint64_t)n_img};
    const uint64_t strides[3] = {(uint64_t)C * ES, (uint64_t)C * ES * VNQA_IMG_W, (uint64_t)C * ES * VNQA_IMG_PIX};
    const uint32_t box[4] = {(uint32_t)W::CH, VNQA_IMG_W, 2, (uint32_t)W::IMG};
    return make_tmap(m, base, Op<T>::HALF, 4, dims, strides, box, true);
  };
  if (make(&tmA, dy, cout) || make(&tmB, x, cin)) return -3;
  const int rc = ncta == 2 ? launch_wgrad<2, T>(tmA, tmB, wp, sms, st) : launch_wgrad<1, T>(tmA, tmB, wp, sms, st);
  if (rc) return rc;
  if (taps != 1) {
    const long long total_w = (long long)cout * cin * taps;
    const int blocks = (int)((total_w + 255) / 256 < sms * 8 ? (total_w + 255) / 256 : sms * 8);
    wgrad_unpack_kernel<<<blocks, 256, 0, st>>>(wp.out, dw, cout, cin, taps);
    VNQA_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace vnqa

using namespace vnqa;

// Workspace of vnqa_conv_wgrad in bytes; *zero_bytes (optional) = size of its leading turn-counter region, which must
// be zero on entry (the kernel leaves it zero again, so a persistent workspace is zeroed once).
extern "C" int64_t vnqa_conv_wgrad_ws_bytes(int32_t cout, int32_t cin, int32_t taps, int64_t* zero_bytes) {
  if (cout <= 0 || cin <= 0 || (taps != 1 && taps != 9)) return -1;
  const WgShape sh = wg_shape(cout, cin, taps);
  if (zero_bytes) *zero_bytes = sh.counter_bytes;
  return sh.counter_bytes + (taps == 1 ? 0 : (long long)taps * cout * cin * 4);
}

// dw is OVERWRITTEN in PyTorch's (cout, cin, kh, kw) layout (no zero-fill needed).
// conv (flat = 0): dy [n_img][130][cout], x [n_img][130][cin].   flat = 1: dy [n_rows][cout], x [n_rows][cin].
extern "C" int vnqa_conv_wgrad(const void* dy, const void* x, float* dw, void* ws, int32_t n_img, int32_t cout,
                               int32_t cin, int32_t taps, int32_t flat, int32_t act_dtype, float grad_unscale,
                               void* stream) {
  VNQA_REQUIRE(dy && x && dw && ws, "vnqa_conv_wgrad: null pointer");
  VNQA_REQUIRE(taps == 1 || taps == 9, "vnqa_conv_wgrad: taps must be 1 or 9");
  VNQA_REQUIRE(!(flat && taps != 1), "vnqa_conv_wgrad: flat GEMM has one tap");
  VNQA_REQUIRE(n_img > 0, "vnqa_conv_wgrad: empty batch");
  VNQA_REQUIRE(act_dtype == VNQA_ACT_TF32 || act_dtype == VNQA_ACT_FP16, "vnqa_conv_wgrad: unknown act_dtype %d", act_dtype);
  return act_dtype == VNQA_ACT_FP16
             ? conv_wgrad_typed<__half>(dy, x, dw, ws, n_img, cout, cin, taps, flat, grad_unscale, as_stream(stream))
             : conv_wgrad_typed<float>(dy, x, dw, ws, n_img, cout, cin, taps, flat, grad_unscale, as_stream(stream));
}
