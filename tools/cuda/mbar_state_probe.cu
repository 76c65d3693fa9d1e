// Bring-up aid: prints the raw 64-bit word of an mbarrier after known operations, to decode the barrier dumps that
// the -DVNQA_DEBUG_MBAR build of libvnqa_b200 writes when a wait stalls (csrc/ptx.cuh, tools/hangwatch.py).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/mbar_state_probe tools/cuda/mbar_state_probe.cu
#include <cstdint>
#include <cstdio>
#include "../../videonavqa_b200/csrc/ptx.cuh"
using namespace vnqa;

__device__ unsigned long long raw(uint64_t* b) {
  unsigned long long w;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(w) : "r"(smem_u32(b)));
  return w;
}
__global__ void probe(unsigned long long* out) {
  __shared__ uint64_t bar;
  if (threadIdx.x != 0) return;
  int n = 0;
  mbar_init(&bar, 2);            out[n++] = raw(&bar);   // init count 2
  mbar_arrive(&bar);             out[n++] = raw(&bar);   // 1 pending
  mbar_expect_tx(&bar, 1024);    out[n++] = raw(&bar);   // arrive + expect 1024 -> pending 0, tx 1024 (phase open)
  asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1024) : "memory");
  out[n++] = raw(&bar);                                   // phase 0 complete -> phase 1, pending 2
  mbar_arrive(&bar);             out[n++] = raw(&bar);
  mbar_arrive(&bar);             out[n++] = raw(&bar);   // phase 1 complete
  asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512) : "memory");
  out[n++] = raw(&bar);                                   // tx = -512 (bytes before expect)
  mbar_init(&bar, 8);            out[n++] = raw(&bar);
  mbar_init(&bar, 1);            out[n++] = raw(&bar);
}
int main() {
  unsigned long long* d;
  cudaMalloc(&d, 16 * 8);
  cudaMemset(d, 0, 16 * 8);
  probe<<<1, 32>>>(d);
  unsigned long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* what[] = {"init(2)", "arrive", "arrive.expect_tx(1024)", "complete_tx(1024) -> phase done", "arrive",
                        "arrive -> phase done", "complete_tx(512) early", "init(8)", "init(1)"};
  for (int i = 0; i < 9; ++i) printf("%-34s %016llx\n", what[i], h[i]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
