// dfma_bench.cu — the FP64 pipe's real ceiling in CYCLES per DFMA and SM sub-partition (clock64 inside the kernel, so SM clock
// changes do not matter), for the operand patterns of the fused kernel and for 1, 2, 4 warps per sub-partition:
//   chain     x[i] = fma(x[i], a, b)              (two operands shared by every instruction: the usual "peak" stream)
//   outer4x4  acc[i][j] += y[i] * z[j]            (real 4x4 outer product: the colour trace of loop_fused_kernel, 16 accumulators)
//   cplx      acc[i][j] += conj(y[i]) * z[j]      (complex 4x4 outer product, 32 accumulators, 4 DFMA per entry: the exact pattern)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o dfma_bench dfma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int PAT>
__global__ void __launch_bounds__(512, 1) k(double *out, const double *in, int iters, long long *cyc) {
  double acc[32], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 32; i++) acc[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 8; i++) { y[i] = in[threadIdx.x + i]; z[i] = in[threadIdx.x + 8 + i]; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (PAT == 0) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], y[0], z[0]);
    } else if (PAT == 1) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i * 4 + j] = fma(y[i], z[j + (r & 1) * 4], acc[i * 4 + j]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double re = acc[2 * (i * 4 + j)], im = acc[2 * (i * 4 + j) + 1];
          re = fma(y[2 * i], z[2 * j], re);
          re = fma(y[2 * i + 1], z[2 * j + 1], re);
          im = fma(y[2 * i], z[2 * j + 1], im);
          im = fma(-y[2 * i + 1], z[2 * j], im);
          acc[2 * (i * 4 + j)] = re;
          acc[2 * (i * 4 + j) + 1] = im;
        }
    }
    // keep the operands changing without FP64 work: swap two of them (register moves are free next to DFMAs)
    const double t = y[it & 1 ? 0 : 1];
    y[it & 1 ? 0 : 1] = z[3];
    z[3] = t;
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PAT> static void run(const char *name, int sms, double *out, double *in, long long *cyc_d) {
  const int iters = 4000;
  for (int warps : {4, 8, 16}) {
    k<PAT><<<sms, warps * 32>>>(out, in, iters, cyc_d);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    k<PAT><<<sms, warps * 32>>>(out, in, iters, cyc_d);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    long long c[256]; CK(cudaMemcpy(c, cyc_d, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < sms; i++) mean += c[i]; mean /= sms;
    const double dfma_per_smsp = (double)iters * 64 * (warps / 4.0);
    printf(", \"%s_%dwarps\": {\"cycles_per_dfma_per_smsp\": %.3f, \"tflops\": %.2f, \"mhz\": %.0f}", name, warps, mean / dfma_per_smsp,
           (double)sms * warps * 32 * iters * 64 * 2 / ms / 1e9, mean / ms / 1e3);
  }
}

// sustained: the same stream for `seconds`; rate and SM clock over the second half (the board's power cap pulls the clock
// down under a long FP64 load: the denominator for a kernel timed inside a long step)
template <int PAT> static void sustained(const char *name, int sms, double *out, double *in, long long *cyc_d, double seconds) {
  const int iters = 40000, warps = 8;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  double elapsed = 0, tf = 0, mhz = 0; int n = 0;
  while (elapsed < seconds) {
    CK(cudaEventRecord(a));
    k<PAT><<<sms, warps * 32>>>(out, in, iters, cyc_d);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    elapsed += ms * 1e-3;
    if (elapsed > seconds / 2) {
      long long c[256]; CK(cudaMemcpy(c, cyc_d, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      double mean = 0; for (int i = 0; i < sms; i++) mean += c[i]; mean /= sms;
      tf += (double)sms * warps * 32 * iters * 64 * 2 / ms / 1e9; mhz += mean / ms / 1e3; n++;
    }
  }
  printf(", \"%s_sustained_%.0fs\": {\"tflops\": %.2f, \"mhz\": %.0f}", name, seconds, tf / n, mhz / n);
}

int main(int argc, char **argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double *out, *in; long long *cyc; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&in, 8192)); CK(cudaMemset(in, 0, 8192)); CK(cudaMalloc(&cyc, 8 * 256));
  printf("{\"gpu\": \"%s\"", prop.name);
  run<0>("chain", sms, out, in, cyc);
  run<1>("outer4x4", sms, out, in, cyc);
  run<2>("cplx_outer4x4", sms, out, in, cyc);
  if (argc > 1) {
    const double sec = atof(argv[1]);
    sustained<0>("chain", sms, out, in, cyc, sec);
    sustained<2>("cplx_outer4x4", sms, out, in, cyc, sec);
  }
  printf("}\n");
  return 0;
}
