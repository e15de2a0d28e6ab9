// microbench.cu — hardware facts the kernel design depends on, measured on the B200 (prints one JSON object):
//   dfma_tflops        FP64 FMA pipe, 8 independent chains per thread
//   dmma_tflops        mma.sync.m8n8k4.f64 (DMMA) alone
//   mixed_*            both instruction streams interleaved in the same warps (do the pipes overlap?)
//   hbm_read_gbs       128-bit streaming read of a 4 GiB buffer (sum-reduced so nothing is elided)
//   l2_read_gbs        same kernel over a 64 MiB buffer (L2 resident)
//   copy_gbs           128-bit copy, read+write bytes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

// per iteration: 8 DMMA (8*256 MAC per warp) + NF*8 DFMA per thread
template <int NF>
__global__ void __launch_bounds__(256) mixed_kernel(double *out, int iters, double a, double b) {
  double c[8][2], x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = c[i][1] = threadIdx.x + i; x[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int f = 0; f < NF; f++) x[(i + f) & 7] = fma(x[(i + f) & 7], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + x[i];
  if (s == 12345.678) out[0] = s;
}

// per DFMA, NI independent integer instructions (IMAD chain on separate registers): does non-FP64 issue steal FP64 slots?
template <int NI, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) dfma_int_kernel(double *out, int iters, double a, double b, int ia) {
  double x[8];
  int y[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x[i] = fma(x[i], a, b);
#pragma unroll
      for (int f = 0; f < NI; f++) y[(i + f) & 7] = y[(i + f) & 7] * ia + it;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i] + y[i];
  if (s == 12345.678) out[0] = s;
}

// DFMA whose three source operands are all distinct registers (no constant, no operand reuse): acc[i] += y[j] * z[k]
// NY y-values x NZ z-values -> NY*NZ accumulators, as in the colour trace of the fused kernel (4 x 4)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) dfma3_kernel(double *out, int iters, const double *in) {
  double y[4], z[4], acc[16];
#pragma unroll
  for (int i = 0; i < 4; i++) { y[i] = in[threadIdx.x + i]; z[i] = in[threadIdx.x + 4 + i]; }
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int k = 0; k < 4; k++) acc[j * 4 + k] = fma(y[j], z[k], acc[j * 4 + k]);
    // keep y, z changing so that the compiler cannot hoist anything (2 extra FP64 ops per 16)
    y[it & 3] += 1e-9;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) read_kernel(const double2 *__restrict__ in, size_t n, double *out) {
  double s = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride), d = __ldg(in + i + 3 * stride);
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n; i += stride) { double2 a = __ldg(in + i); s += a.x + a.y; }
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ in, double2 *__restrict__ outp, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride), d = __ldg(in + i + 3 * stride);
    outp[i] = a; outp[i + stride] = b; outp[i + 2 * stride] = c; outp[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) outp[i] = __ldg(in + i);
}

template <typename L> static float time_ms(L launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, 64));
  const int iters = 20000, blocks = sms * 8, threads = 256;
  const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
  float t;
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  t = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  printf(", \"dfma_tflops\": %.2f", nthreads * iters * 8.0 * 2 / t / 1e9);
  t = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  printf(", \"dmma_tflops\": %.2f", nwarps * iters * 8.0 * 256 * 2 / t / 1e9);
  t = time_ms([&] { mixed_kernel<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  printf(", \"mixed1_tflops\": %.2f, \"mixed1_ms\": %.3f", (nwarps * iters * 8.0 * 256 * 2 + nthreads * iters * 8.0 * 2) / t / 1e9, t);
  t = time_ms([&] { mixed_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  printf(", \"mixed4_tflops\": %.2f, \"mixed4_ms\": %.3f", (nwarps * iters * 8.0 * 256 * 2 + nthreads * iters * 32.0 * 2) / t / 1e9, t);
  t = time_ms([&] { mixed_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  printf(", \"mixed8_tflops\": %.2f, \"mixed8_ms\": %.3f", (nwarps * iters * 8.0 * 256 * 2 + nthreads * iters * 64.0 * 2) / t / 1e9, t);

  // 8 warps per SM (2 per sub-partition), as in the fused loop kernel
  {
    const double n8 = (double)sms * 256;
    t = time_ms([&] { dfma_int_kernel<0, 8><<<sms, 256>>>(out, iters, 1.0000001, 1e-9, 3); }, 5);
    printf(", \"dfma_8warps_tflops\": %.2f", n8 * iters * 8.0 * 2 / t / 1e9);
    t = time_ms([&] { dfma_int_kernel<1, 8><<<sms, 256>>>(out, iters, 1.0000001, 1e-9, 3); }, 5);
    printf(", \"dfma_8warps_1int_tflops\": %.2f", n8 * iters * 8.0 * 2 / t / 1e9);
    t = time_ms([&] { dfma_int_kernel<2, 8><<<sms, 256>>>(out, iters, 1.0000001, 1e-9, 3); }, 5);
    printf(", \"dfma_8warps_2int_tflops\": %.2f", n8 * iters * 8.0 * 2 / t / 1e9);
    t = time_ms([&] { dfma_int_kernel<1, 16><<<sms, 512>>>(out, iters, 1.0000001, 1e-9, 3); }, 5);
    printf(", \"dfma_16warps_1int_tflops\": %.2f", n8 * 2 * iters * 8.0 * 2 / t / 1e9);
  }
  {
    double *zin; CK(cudaMalloc(&zin, 4096 * 8)); CK(cudaMemset(zin, 0, 4096 * 8));
    const double n8 = (double)sms * 256;
    t = time_ms([&] { dfma3_kernel<8><<<sms, 256>>>(out, iters, zin); }, 5);
    printf(", \"dfma_3distinct_8warps_tflops\": %.2f", n8 * iters * 16.0 * 2 / t / 1e9);
    t = time_ms([&] { dfma3_kernel<16><<<sms, 512>>>(out, iters, zin); }, 5);
    printf(", \"dfma_3distinct_16warps_tflops\": %.2f", n8 * 2 * iters * 16.0 * 2 / t / 1e9);
  }
  const size_t big = (size_t)4 << 30, small = (size_t)64 << 20;
  double2 *buf, *buf2; CK(cudaMalloc(&buf, big)); CK(cudaMalloc(&buf2, big));
  CK(cudaMemset(buf, 0, big)); CK(cudaMemset(buf2, 0, big));
  for (int occ : {4, 8, 16}) {
    t = time_ms([&] { read_kernel<<<sms * occ, 256>>>(buf, big / 16, out); }, 5);
    printf(", \"hbm_read_gbs_occ%d\": %.1f", occ, big / t / 1e6);
  }
  t = time_ms([&] { read_kernel<<<sms * 8, 256>>>(buf, small / 16, out); }, 20);
  printf(", \"l2_read_gbs\": %.1f", small / t / 1e6);
  t = time_ms([&] { copy_kernel<<<sms * 8, 256>>>(buf, buf2, big / 16); }, 5);
  printf(", \"copy_gbs\": %.1f", 2.0 * big / t / 1e6);
  printf("}\n");
  return 0;
}
