// tma.cuh — bulk-TMA (cp.async.bulk) + mbarrier primitives for the per-stage kernels (displace.cu, contract.cu).
//
// sm_100a only: the copies are linear bulk copies (no tensor map) because every operand of these kernels is a
// contiguous run of whole sites in the site-major layout (include/mugiq_b200.h); SASS shows them as UBLKCP.
// The fused kernel keeps its own predicated variants (fused_kernel.cu), tuned for its FP64-issue-bound loop.
#pragma once
#include "common.cuh"

namespace mugiq_b200 {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TMA_DONE;\n"
      "bra TMA_WAIT;\n"
      "TMA_DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, completion counted in bytes on `bar`; addresses and size multiples of 16 B
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy writes to shared memory visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename F> __device__ __forceinline__ Cplx<F> lds_c(const char *p) {
  using V = typename vec2_of<F>::type;
  const V v = *reinterpret_cast<const V *>(p);
  return make_c<F>(v.x, v.y);
}
template <typename F> __device__ __forceinline__ void sts_c(char *p, const Cplx<F> c) {
  using V = typename vec2_of<F>::type;
  V v;
  v.x = c.re;
  v.y = c.im;
  *reinterpret_cast<V *>(p) = v;
}

}  // namespace tma
}  // namespace mugiq_b200
