// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM).
// No CUTLASS dependency; bit layouts follow the PTX ISA "tcgen05" matrix/instruction descriptor tables.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace vnqa {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Round-to-nearest fp32 -> tf32 (10-bit mantissa), returned as fp32 bits.  The tensor core TRUNCATES fp32
// operands of kind::tf32 (measured: bit-exact against a truncating reference), which is a biased error;
// every tensor that feeds an MMA is therefore rounded once, here, by the kernel that produces it.
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Exactly one lane of a fully converged warp returns true.  Using elect.sync (instead of `lane == 0`) lets ptxas
// see that the guarded region has a single active thread, so tcgen05.mma / TMA operands move to uniform registers
// with one R2UR instead of a per-lane "waterfall" loop around every instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Blocking wait.  try_wait suspends the thread for a hardware time slice, so the loop is not a hot spin.  Release
// builds wait without a bound: a legitimately long wait (SM time-slicing, a co-resident collective, a profiler or
// sanitizer replay) must not abort training.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// -DVNQA_DEBUG_MBAR (libvnqa_b200_dbg.so, pipeline bring-up only): a wait that lasts longer than 2 s of %globaltimer
// writes a record into a host-mapped buffer installed with vnqa_debug_install() - who waits (block, thread, role tag,
// loop counters) plus the raw 64-bit state of every mbarrier of the CTA - and keeps waiting, so that a host-side
// monitor (tools/hangwatch.py) can print what a wedged kernel is waiting for without any CUDA call.
struct DbgCtx {
  unsigned long long* buf;   // host-mapped, VNQA_DBG_SLOTS records of VNQA_DBG_REC words; nullptr = off
  const uint64_t* bars;      // first mbarrier of this CTA
  int nbars;
};
constexpr int VNQA_DBG_REC = 32;
constexpr int VNQA_DBG_SLOTS = 148 * 8;
#ifdef VNQA_DEBUG_MBAR
static __device__ __noinline__ void mbar_dbg_record(const DbgCtx& d, uint64_t* bar, uint32_t parity, int tag, long long info,
                                             unsigned rounds) {
  if (d.buf == nullptr) return;
  const int slot = (blockIdx.x * 8 + (threadIdx.x >> 5)) % VNQA_DBG_SLOTS;
  volatile unsigned long long* r = d.buf + (size_t)slot * VNQA_DBG_REC;
  uint32_t smid, crank;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  r[1] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
  r[2] = ((unsigned long long)tag << 32) | (parity << 16) | (unsigned)(bar - d.bars);
  r[3] = (unsigned long long)info;
  r[4] = ((unsigned long long)smid << 32) | (crank << 16) | rounds;
  for (int i = 0; i < d.nbars && i < VNQA_DBG_REC - 6; ++i) {
    unsigned long long w;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(w) : "r"(smem_u32(d.bars + i)));
    r[6 + i] = w;
  }
  __threadfence_system();
  r[0] = 0x564E5141ull;   // "VNQA": record valid
}
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, int tag, long long info, const DbgCtx& d) {
  uint64_t t0 = 0;
  uint32_t spins = 0, rounds = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull) {
        if ((threadIdx.x & 31) == 0) mbar_dbg_record(d, bar, parity, tag, info, ++rounds);
        t0 = now;
      }
    }
  }
}
#define MBAR_WAIT(bar, parity, tag, info) mbar_wait_dbg(bar, parity, tag, (long long)(info), dbg)
#else
#define MBAR_WAIT(bar, parity, tag, info) mbar_wait(bar, parity)
#endif

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 operands, FP32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64 bit):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100)   [49,52) base offset   [61,64) layout: 0 none, 2 = 128B swizzle
// K-major, 128B swizzle: rows are 128 B apart, 8-row groups are SBO = 1024 B apart, LBO unused (1).
__host__ __device__ constexpr uint64_t umma_desc_base_sw128(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}
// MN-major fp32/tf32 operands must use the "128B swizzle with 32B atoms" layout (type 1): 32-byte chunks of a
// 128-byte row are XORed with (row % 4); K rows come in groups of 4 (SBO = 512 B when rows are dense), the next
// 32 MN elements are LBO bytes away.  The matching TMA mode is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__host__ __device__ constexpr uint64_t umma_desc_base_sw128_32b(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(1) << 61);
}
__device__ __forceinline__ uint64_t umma_desc(uint64_t base, uint32_t smem_addr) {
  return base | uint64_t((smem_addr >> 4) & 0x3FFF);
}
// Instruction descriptor (32 bit) for kind::tf32, FP32 accumulate:
//   [4,6) D format 1 = F32   [7,10) A format 2 = TF32   [10,13) B format 2 = TF32
//   [15] A major (0 = K)   [16] B major (0 = K)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// Instruction descriptor for kind::f16 with F16 A/B (format 0) and F32 D:
//   [4,6) D format 1 = F32   [7,10) A format 0 = F16   [10,13) B format 0 = F16
//   [15] A major   [16] B major (0 = K)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}
// MN-major, 128B swizzle, 16-bit elements: layout type 2, LBO = bytes to the next 64 MN elements, SBO = bytes to the
// next 8 K rows (for 16-bit types the ordinary 128-byte swizzle IS the canonical MN-major layout)
__host__ __device__ constexpr uint64_t umma_desc_base_mn_sw128(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}
// D (+)= A * B with fp16 operands from shared memory, fp32 accumulate in TMEM; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace vnqa

// ---------------------------------------------------------------- 2-CTA (cta_group::2) variants
// A CTA pair (cluster of 2 on one TPC) executes ONE UMMA of M = 256: each CTA supplies its own 128 rows of A and
// half of the N rows of B from its own shared memory (same offsets in both CTAs) and receives its 128 x N half of the
// accumulator in its own TMEM.  Only the leader (cluster rank 0) issues the MMA; both CTAs issue TMA loads that
// signal the LEADER's mbarrier (the pair bit 24 of the shared::cluster address is cleared).
namespace vnqa {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// arrive + expect `bytes` on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_expect_tx_remote(uint64_t* bar, uint32_t bytes, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %2;\n"
      "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [ra], %1;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(cta)
      : "memory");
}

// ---- cluster launch control (dynamic tile scheduling in hardware): a running cluster asks to CANCEL a cluster of the
// same grid that has not been launched yet and, when that succeeds, does its work instead.  The 16-byte response lands
// in shared memory (in every CTA of the cluster with MULTICAST) and completes 16 bytes on the mbarrier.
template <bool MULTICAST>
__device__ __forceinline__ void clc_try_cancel(void* resp16, uint64_t* bar) {
  if (MULTICAST)
    asm volatile(
        "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 "
        "[%0], [%1];" ::"r"(smem_u32(resp16)), "r"(smem_u32(bar)) : "memory");
  else
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(
                     smem_u32(resp16)), "r"(smem_u32(bar)) : "memory");
}
// first CTA id (x) of the cancelled cluster, or -1 when nothing was left to cancel
__device__ __forceinline__ int clc_response_ctaid_x(const void* resp16) {
  uint32_t valid, x;
  asm volatile(
      "{\n"
      ".reg .pred p1;\n"
      ".reg .b128 r;\n"
      "ld.shared.b128 r, [%2];\n"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n"
      "selp.u32 %0, 1, 0, p1;\n"
      "mov.u32 %1, 0;\n"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, r;\n"
      "}\n"
      : "=r"(valid), "=r"(x)
      : "r"(smem_u32(resp16))
      : "memory");
  return valid ? (int)x : -1;
}

__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint64_t* leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* m, uint64_t* leader_bar, void* dst, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have retired) on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

}  // namespace vnqa
