// tma_bench.cu — how fast can one CTA per SM stream many SMALL contiguous chunks into shared memory?
// Pattern of the fused loop kernel: per "stage" NCOPY chunks of CHUNK bytes (scattered 192-B-multiple offsets in a
// large buffer), ring of 4 stages, consumers only wait.  Variants:
//   mode 0: cp.async.bulk, one lane issues all copies            mode 1: cp.async.bulk, 32 lanes of warp 0 issue
//   mode 2: cp.async.bulk, lane 0 of each of 8 warps issues NCOPY/8    mode 3: LDGSTS (cp.async 16 B), all 256 threads
// Prints GB/s and bytes/clk/SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void *d, const void *s, uint32_t n, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void ldgsts16(void *d, const void *s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d)), "l"(s) : "memory"); }
__device__ __forceinline__ void ldgsts_arrive(uint64_t *b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(b)) : "memory"); }

constexpr int S = 4;
__global__ void __launch_bounds__(256, 1) k(const char *buf, size_t bufbytes, int niter, int ncopy, int chunk, int mode, double *sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *full = (uint64_t *)smem, *empty = full + 8;
  char *stages = (char *)smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = ncopy * chunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], mode == 3 ? 256 : (mode == 2 ? 8 : 1)); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // chunk c of iteration n of this CTA: pseudo-random 192*64-aligned offset
  auto src = [&](int n, int c) -> const char * {
    size_t h = ((size_t)blockIdx.x * 7919u + (size_t)n * 104729u + (size_t)c * 1299709u) * 12288u;
    return buf + (h % (bufbytes - 65536)) / 12288 * 12288;
  };
  auto produce = [&](int n) {  // all threads call; roles inside
    const int s = n % S;
    char *dst = stages + (size_t)s * stage_bytes;
    if (mode == 0) {
      if (threadIdx.x == 0) { mbar_expect(&full[s], stage_bytes); for (int c = 0; c < ncopy; c++) bulk(dst + c * chunk, src(n, c), chunk, &full[s]); }
    } else if (mode == 1) {
      if (warp == 0) { if (lane == 0) mbar_expect(&full[s], stage_bytes); __syncwarp(); for (int c = lane; c < ncopy; c += 32) bulk(dst + c * chunk, src(n, c), chunk, &full[s]); }
    } else if (mode == 2) {
      if (lane == 0) { int per = (ncopy + 7) / 8, c0 = warp * per, c1 = min(ncopy, c0 + per); mbar_expect(&full[s], (c1 > c0 ? c1 - c0 : 0) * chunk); for (int c = c0; c < c1; c++) bulk(dst + c * chunk, src(n, c), chunk, &full[s]); }
    } else {
      const int per = chunk / 16;
      for (int i = threadIdx.x; i < ncopy * per; i += 256) { int c = i / per, o = (i % per) * 16; ldgsts16(dst + c * chunk + o, src(n, c) + o); }
      ldgsts_arrive(&full[s]);
    }
  };
  for (int n = 0; n < S - 1 && n < niter; n++) produce(n);
  double acc = 0;
  for (int n = 0; n < niter; n++) {
    const int m = n + S - 1;
    if (m < niter) {
      if (m >= S) mbar_wait(&empty[m % S], (uint32_t)(((m / S) & 1) ^ 1));
      produce(m);
    }
    const int s = n % S;
    mbar_wait(&full[s], (uint32_t)((n / S) & 1));
    acc += *(const double *)(stages + (size_t)s * stage_bytes + (threadIdx.x * 16) % stage_bytes);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (acc == 1.2345) sink[0] = acc;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const size_t bufbytes = (size_t)2 << 30;
  char *buf; CK(cudaMalloc(&buf, bufbytes)); CK(cudaMemset(buf, 1, bufbytes));
  double *sink; CK(cudaMalloc(&sink, 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("[");
  bool first = true;
  const int cfgs[][2] = {{24, 1536}, {8, 6144}, {4, 12288}, {2, 24576}, {48, 768}};
  for (auto &cfg : cfgs)
    for (int mode = 0; mode < 4; mode++) {
      const int ncopy = cfg[0], chunk = cfg[1], niter = 400;
      const size_t smem = 1024 + (size_t)S * ncopy * chunk;
      k<<<sms, 256, smem>>>(buf, bufbytes, 20, ncopy, chunk, mode, sink);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a));
      k<<<sms, 256, smem>>>(buf, bufbytes, niter, ncopy, chunk, mode, sink);
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      const double bytes = (double)sms * niter * ncopy * chunk;
      printf("%s\n{\"ncopy\": %d, \"chunk\": %d, \"mode\": %d, \"GBps\": %.1f, \"B_per_clk_per_SM\": %.1f, \"us_per_stage\": %.2f}", first ? "" : ",", ncopy, chunk, mode,
             bytes / ms / 1e6, bytes / sms / (ms * 1e-3 * clk * 1e3), ms * 1e3 / niter);
      first = false;
    }
  printf("\n]\n");
  return 0;
}
