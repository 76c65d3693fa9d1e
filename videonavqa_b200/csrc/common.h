// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/vnqa_b200.h"

namespace vnqa {

void set_error(const char* fmt, ...);

#define VNQA_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::vnqa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)

#define VNQA_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      ::vnqa::set_error(__VA_ARGS__); \
      return -1;                     \
    }                                \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// Tiled tensor map (fp32 or fp16 elements) with 128-byte swizzle and zero out-of-bounds fill.  dims / box innermost
// first; strides in bytes for dims 1..rank-1.  atom32 selects the "128B swizzle, 32B atoms" layout that MN-major
// fp32/tf32 MMA operands need (16-bit operands use the ordinary 128B swizzle in both majors).  Encoded maps are kept
// in a small direct-mapped cache keyed by every argument: a training loop presents the same (pointer, shape) pairs
// step after step (PyTorch's caching allocator returns the same blocks), so cuTensorMapEncodeTiled is rarely re-run.
int make_tmap(CUtensorMap* m, const void* base, bool half, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, bool atom32 = false);

// host-mapped diagnostics buffer installed by vnqa_debug_install (nullptr unless a debug session set one)
unsigned long long* debug_buffer();

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace vnqa
