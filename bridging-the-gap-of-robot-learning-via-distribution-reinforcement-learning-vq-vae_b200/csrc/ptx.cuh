// vqb200 -- thin PTX wrappers shared by the sm_100a kernels: mbarrier, bulk-TMA copies, proxy fences.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vqb200 {
namespace ptx {

constexpr unsigned SPIN_LIMIT = 1u << 24;   // bounded waits: a protocol bug traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" :: "r"(bar), "r"(bytes) : "memory");
}
// Plain try_wait (the hardware parks the thread for an implementation-defined, short time).  A variant with an explicit
// suspend-time hint (20 us) was measured: ~2 % faster, but with it the streaming assignment kernel hung reproducibly at
// K >= 32 768 right after the EMA kernels (producer and MMA-issuer threads parked inside try_wait and never resumed,
// every mbarrier idle) -- removed.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// non-blocking probe (mbarrier.test_wait never suspends the thread)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// BACKOFF_NS > 0: sleep between probes.  Roles that wait long and are not on the critical path (producer, converter, MMA
// issuer) must not burn issue slots: their spin loops were ~18 % of all executed instructions of the assignment kernel.
template <int BACKOFF_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code, uint32_t dump_base = 0, int dump_n = 0) {
  unsigned spins = 0;
  long long t0 = 0;
  while (!mbar_try(bar, parity)) {
    if (BACKOFF_NS > 0) __nanosleep(BACKOFF_NS);
    if (spins == 0) t0 = clock64();
    ++spins;
    if (spins == SPIN_LIMIT / 4 && (threadIdx.x & 31) == 0)      // report every stuck role before the first one traps
    {
      printf("vqb200: mbarrier wait %d stuck for %lld cycles (block %d, warp %d, parity %u)\n", code, clock64() - t0,
             (int)blockIdx.x, (int)(threadIdx.x >> 5), parity);
      for (int i = 0; i < dump_n; ++i) {
        unsigned long long w;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(w) : "r"(dump_base + 8 * i));
        printf("vqb200:   block %d bar[%d] = %016llx\n", (int)blockIdx.x, i, w);
      }
    }
    if (spins > SPIN_LIMIT) { if (err) atomicExch(err, code); __trap(); }
  }
}
// 1-D bulk-TMA copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 1-D bulk-TMA copy shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace ptx
}  // namespace vqb200
