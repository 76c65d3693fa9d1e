// K4: weight gradients on tcgen05 tensor cores (autograd of the convolutions / Linear of
// models/film_attn_pt_stem.py:211,219,224,244).
//
//   dW[tap][co][ci] = sum_{pixel rows r} dY[r][co] * X[r (+) tap][ci]
//
// The reduction runs over PIXELS, and both operands are stored pixel-major ([row][channel], channel contiguous),
// i.e. "MN-major" for the MMA: no transposed copies are made - the UMMA descriptors are built MN-major
// (instruction-descriptor bits 15/16, smem layout "128B swizzle / 32B atoms", the only one tf32 accepts
// MN-major) over TMA boxes {32 channels x 104 pixel rows}
// (104 = 4 images x 2 image rows x 13 = 13 MMA K-steps of 8).  Tap shifts and the zero halo are again TMA box
// coordinates / out-of-bounds fill.  fp32 accumulators in TMEM (double-buffered), 2 smem stages of 104 KB.
//
// Two variants (template NCTA):
//   NCTA = 2 (Cout and Cin multiples of 256): a CTA pair executes tcgen05.mma.cta_group::2 on a 256 (co) x 256 (ci)
//     tile - each CTA stages its own 128 output channels of dY and its own 128 input channels of X, halving the
//     operand bytes per MMA flop with respect to the single-CTA 128 x 128 tile;
//   NCTA = 1: 128 x 128 tiles (small / odd channel counts, the Linear of :244).
// Split-K over pixel chunks, SPLIT-MAJOR tile order: the co-resident CTAs always work on the same
// pixel range (sized to stay L2-resident), so dY / X are read from DRAM about once instead of once per tile wave.
// Partial tiles are reduced with fp32 red.global.add, transposed through shared memory so that every warp-level
// reduction hits 128 contiguous bytes of a [tap][Cout][Cin] buffer; a final pass writes PyTorch's (Cout,Cin,kh,kw).
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace vnqa {

constexpr int WG_M = 128;                        // output channels per CTA (TMEM lanes)
constexpr int WG_NB = 128;                       // input channels STAGED per CTA
constexpr int WG_KP = 104;                       // pixel rows per stage
constexpr int WG_BOX_BYTES = WG_KP * 128;        // 13312: one {32 ch x 104 rows} box
constexpr int WG_A_BYTES = (WG_M / 32) * WG_BOX_BYTES;   // 53248
constexpr int WG_B_BYTES = (WG_NB / 32) * WG_BOX_BYTES;  // 53248
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;  // 106496
constexpr int WG_STAGES = 2;
constexpr int WG_SCRATCH = 4 * 32 * 33 * 4;      // per-epilogue-warp 32x33 transpose tile
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_SCRATCH + 1024 + 256;
constexpr int WG_THREADS = 256;

struct WParams {
  int taps, m_tiles, n_tiles, splits, n_chunks, chunks_per_split, flat;
  int cout, cin;          // real sizes (tiles beyond are skipped on store)
  float* out;             // [taps][cout][cin] accumulation buffer (== dw when taps == 1)
  unsigned long long* dbg;   // VNQA_DEBUG_MBAR builds only
};

// m_tiles counts 128-row co tiles, n_tiles counts (128*NCTA)-wide ci tiles.
template <int NCTA>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WParams p) {
  constexpr int N_INSTR = WG_NB * NCTA;          // MMA N = accumulator columns per CTA
  constexpr int TMEM_COLS = 2 * N_INSTR;         // two accumulator stages (256 / 512 columns)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* scratch = reinterpret_cast<float*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES + WG_SCRATCH);
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
          uint8_t* sa = smem + stage * WG_STAGE_BYTES;
          uint8_t* sb = sa + WG_A_BYTES;
          int a1, a2, a3, b1, b2, b3;
          if (p.flat) { a1 = c * WG_KP; a2 = 0; a3 = 0; b1 = a1; b2 = 0; b3 = 0; }
          else { a1 = 0; a2 = (c % 5) * 2; a3 = (c / 5) * 4; b1 = dx; b2 = a2 + dy; b3 = a3; }
          if (NCTA == 1) {
            mbar_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < WG_M / 32; ++i)
              tma_load_4d(&tmA, &full_bar[stage], sa + i * WG_BOX_BYTES, mt * WG_M + i * 32, a1, a2, a3);
#pragma unroll
            for (int i = 0; i < WG_NB / 32; ++i)
              tma_load_4d(&tmB, &full_bar[stage], sb + i * WG_BOX_BYTES, ci0 + i * 32, b1, b2, b3);
          } else {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * WG_STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < WG_M / 32; ++i)
              tma_load_4d_2sm(&tmA, &full_bar[stage], sa + i * WG_BOX_BYTES, mt * WG_M + i * 32, a1, a2, a3);
#pragma unroll
            for (int i = 0; i < WG_NB / 32; ++i)
              tma_load_4d_2sm(&tmB, &full_bar[stage], sb + i * WG_BOX_BYTES, ci0 + i * 32, b1, b2, b3);
            if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
          }
        }
        __syncwarp();
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    constexpr uint32_t idesc = umma_idesc_tf32(WG_M * NCTA, N_INSTR, 1, 1);       // both operands MN-major
    constexpr uint64_t dbase = umma_desc_base_sw128_32b(WG_BOX_BYTES, 512);       // LBO = next 32 channels, SBO = next 4 rows
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
          const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
          const uint64_t adesc = umma_desc(dbase, sa);
          const uint64_t bdesc = umma_desc(dbase, sa + WG_A_BYTES);
#pragma unroll
          for (int g = 0; g < WG_KP / 8; ++g) {   // 8 pixel rows = 1024 bytes (>>4 = 64) per K-step
            if (NCTA == 1) umma_tf32(d_tmem, adesc + 64 * g, bdesc + 64 * g, idesc, (c > c_begin || g > 0) ? 1u : 0u);
            else umma_tf32_2sm(d_tmem, adesc + 64 * g, bdesc + 64 * g, idesc, (c > c_begin || g > 0) ? 1u : 0u);
          }
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
      const uint32_t t_base = tmem_base + (uint32_t(ew * 32) << 16) + as * N_INSTR;
      float* tile_s = scratch + ew * (32 * 33);
      const int co0 = mt * WG_M + ew * 32;           // this warp owns output channels co0 .. co0+31
#pragma unroll 1
      for (int j0 = 0; j0 < N_INSTR; j0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(t_base + j0, v);
        tmem_ld_wait();
        // lane = co holds 32 consecutive ci: transpose through smem so that a warp reduces 32 consecutive ci of ONE co
#pragma unroll
        for (int j = 0; j < 32; ++j) tile_s[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int ci = nt * N_INSTR + j0 + lane;
        if (ci < p.cin) {
          float* obase = p.out + ((long long)tap * p.cout + co0) * p.cin + ci;
#pragma unroll 8
          for (int r = 0; r < 32; ++r)
            if (co0 + r < p.cout) atomicAdd(obase + (long long)r * p.cin, tile_s[r * 33 + lane]);
        }
        __syncwarp();
      }
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

template <int NCTA>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const WParams& wp, int sms, cudaStream_t st) {
  auto kern = wgrad_kernel<NCTA>;
  VNQA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  const int items = (wp.m_tiles / NCTA) * wp.n_tiles * wp.taps * wp.splits;
  const int groups = sms / NCTA;
  const int grid = (items < groups ? items : groups) * NCTA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = WG_SMEM;
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

}  // namespace vnqa

using namespace vnqa;

// taps == 1: partial tiles are accumulated atomically into dw, which must be zero-initialised by the caller.
// taps == 9: accumulation happens in ws ([9][cout][cin], zero-initialised by the caller); dw is then overwritten
//            in PyTorch (cout, cin, 3, 3) layout.
// conv (flat = 0): dy [n_img][130][cout], x [n_img][130][cin].   flat = 1: dy [n_rows][cout], x [n_rows][cin].
extern "C" int vnqa_conv_wgrad(const float* dy, const float* x, float* dw, float* ws, int32_t n_img, int32_t cout,
                               int32_t cin, int32_t taps, int32_t flat, void* stream) {
  VNQA_REQUIRE(dy && x && dw, "vnqa_conv_wgrad: null pointer");
  VNQA_REQUIRE(taps == 1 || ws != nullptr, "vnqa_conv_wgrad: 3x3 gradients need the [9][cout][cin] workspace");
  VNQA_REQUIRE(taps == 1 || taps == 9, "vnqa_conv_wgrad: taps must be 1 or 9");
  VNQA_REQUIRE(cout % 32 == 0 && cin % 32 == 0, "vnqa_conv_wgrad: channels must be multiples of 32 (%d, %d)", cout, cin);
  VNQA_REQUIRE(!(flat && taps != 1), "vnqa_conv_wgrad: flat GEMM has one tap");
  int ncta = (cout % 256 == 0 && cin % 256 == 0) ? 2 : 1;
  if (const char* e = getenv("VNQA_WGRAD_NCTA")) ncta = (atoi(e) == 1) ? 1 : ncta;   // bring-up switch
  WParams wp{};
  wp.taps = taps;
  wp.m_tiles = (cout + WG_M - 1) / WG_M;
  wp.n_tiles = (cin + WG_NB * ncta - 1) / (WG_NB * ncta);
  wp.flat = flat;
  wp.cout = cout;
  wp.cin = cin;
  wp.n_chunks = flat ? (n_img + WG_KP - 1) / WG_KP : ((n_img + 3) / 4) * 5;
  wp.out = taps == 1 ? dw : ws;
  wp.dbg = debug_buffer();
  const int sms = vnqa_sm_count();
  if (sms <= 0) return -2;
  // split-K: (a) one split's slice of dY and X (chunks * 104 rows * (cout+cin) * 4 B) should stay L2-resident while
  // the waves of that split run, (b) there must be enough tiles to fill the SMs, (c) every split owns >= 1 chunk.
  const int base_items = (wp.m_tiles / ncta) * wp.n_tiles * taps;
  const int groups = sms / ncta;
  const double chunk_bytes = (double)WG_KP * 4.0 * ((double)(cout < 1024 ? cout : 1024) + (double)(cin < 1024 ? cin : 1024));
  int cps = (int)(48.0e6 / chunk_bytes);                       // ~48 MB of the 126 MB L2 per operand pair
  const int cps_fill = (int)((long long)wp.n_chunks * base_items / (2LL * groups));   // >= 2 items per CTA group
  if (cps > cps_fill) cps = cps_fill;
  if (cps < 4) cps = 4;
  if (cps > wp.n_chunks) cps = wp.n_chunks;
  wp.chunks_per_split = cps;
  wp.splits = (wp.n_chunks + cps - 1) / cps;

  CUtensorMap tmA, tmB;
  auto make = [&](CUtensorMap* m, const float* base, int C) -> int {
    if (flat) {
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)n_img, 1, 1};
      const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)C * 4 * n_img, (uint64_t)C * 4 * n_img};
      const uint32_t box[4] = {32, WG_KP, 1, 1};
      return make_tmap_f32(m, base, 4, dims, strides, box, true);
    }
    const uint64_t dims[4] = {(uint64_t)C, VNQA_IMG_W, VNQA_IMG_H, (uint64_t)n_img};
    const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)C * 4 * VNQA_IMG_W, (uint64_t)C * 4 * VNQA_IMG_PIX};
    const uint32_t box[4] = {32, VNQA_IMG_W, 2, 4};
    return make_tmap_f32(m, base, 4, dims, strides, box, true);
  };
  if (make(&tmA, dy, cout) || make(&tmB, x, cin)) return -3;
  const int rc = ncta == 2 ? launch_wgrad<2>(tmA, tmB, wp, sms, as_stream(stream))
                           : launch_wgrad<1>(tmA, tmB, wp, sms, as_stream(stream));
  if (rc) return rc;
  if (taps != 1) {
    const long long total_w = (long long)cout * cin * taps;
    const int blocks = (int)((total_w + 255) / 256 < sms * 8 ? (total_w + 255) / 256 : sms * 8);
    wgrad_unpack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(ws, dw, cout, cin, taps);
    VNQA_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
