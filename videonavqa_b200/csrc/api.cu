// Library-wide C-ABI plumbing: error string, device check, tensor-map encoding.
#include <cstring>
#include <mutex>

#include "common.h"

namespace vnqa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

namespace {
struct TmapKey {
  const void* base;
  uint64_t dims[4], strides[3];
  uint32_t box[4];
  int rank, half, atom32, dev;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapSlot {
  bool valid = false;
  TmapKey key;
  CUtensorMap map;
};
constexpr int TMAP_SLOTS = 256;
TmapSlot g_tmaps[TMAP_SLOTS];
std::mutex g_tmap_mu;
}  // namespace

int make_tmap(CUtensorMap* m, const void* base, bool half, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, bool atom32) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return -3;
  }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.half = half; key.atom32 = atom32;
  cudaGetDevice(&key.dev);
  uint64_t h = reinterpret_cast<uint64_t>(base) * 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i > 0) key.strides[i - 1] = strides_bytes[i - 1];
    h = (h ^ (dims[i] * 0xC2B2AE3D27D4EB4Full + box[i])) * 0x9E3779B97F4A7C15ull;
  }
  const int slot = (int)((h >> 40) % TMAP_SLOTS);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps[slot].valid && g_tmaps[slot].key == key) {
      *m = g_tmaps[slot].map;
      return 0;
    }
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                   const_cast<void*>(base), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (atom32 && !half) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, %s, dims %llu %llu, box %u %u)", (int)r, rank,
              half ? "fp16" : "fp32", (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return -3;
  }
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  g_tmaps[slot].key = key;
  g_tmaps[slot].map = *m;
  g_tmaps[slot].valid = true;
  return 0;
}

static unsigned long long* g_dbg = nullptr;
unsigned long long* debug_buffer() { return g_dbg; }

}  // namespace vnqa

extern "C" int vnqa_debug_install(void* host_mapped) {
  vnqa::g_dbg = static_cast<unsigned long long*>(host_mapped);
  return 0;
}
extern "C" int vnqa_debug_build(void) {
#ifdef VNQA_DEBUG_MBAR
  return 1;
#else
  return 0;
#endif
}

extern "C" int vnqa_abi_version(void) { return VNQA_ABI_VERSION; }
extern "C" const char* vnqa_last_error(void) { return vnqa::g_err; }

extern "C" int vnqa_sm_count(void) {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    vnqa::set_error("cudaGetDevice failed");
    return -2;
  }
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      vnqa::set_error("cudaDeviceGetAttribute failed");
      return -2;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" int vnqa_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  VNQA_CHECK_CUDA(cudaGetDevice(&dev));
  VNQA_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  VNQA_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  VNQA_REQUIRE(major == 10 && minor == 0, "libvnqa_b200 needs an sm_100 device (B200); found sm_%d%d", major, minor);
  return 0;
}
