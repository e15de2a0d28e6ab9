// issue_bench.cu — what does ONE extra instruction of a given class cost a DFMA-bound warp on B200?
// 8 warps per SM (2 per sub-partition, as in loop_fused_kernel), every thread runs 16 independent DFMA chains and, after
// every RATIO DFMAs, one instruction X.  Prints TFLOP/s of the DFMAs alone per (X, RATIO): with cost c (in DFMA slots) the
// rate is peak * RATIO / (RATIO + c).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_bench issue_bench.cu
// (check the SASS: cuobjdump -sass issue_bench | grep -A40 'issue_kernel<X')
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { X_NONE, X_IMAD, X_IADD, X_LOP, X_WAIT, X_LDS128, X_LDS64, X_LDS128U, X_ARRIVE, X_SYNCWARP, X_FFMA, X_SHFL, X_UADD, X_DADD, X_COUNT };
static const char *kNames[] = {"none", "imad", "iadd", "lop3", "mbar_try_wait", "lds128", "lds64", "lds128_plus_uniform", "mbar_arrive_lane0", "syncwarp", "ffma", "shfl", "uniform_imad", "dadd"};


template <int X, int RATIO>
__global__ void __launch_bounds__(256, 1) issue_kernel(double *out, int iters, double a, double b, int ia, int ub) {
  extern __shared__ __align__(16) unsigned char sm[];
  double x[16];
  int y[4];
  float f[4];
  double l0[4], l1[4];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 4; i++) { y[i] = threadIdx.x * (i + 1) + ia; f[i] = threadIdx.x + i; l0[i] = l1[i] = 0; }
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 16;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(sm) + 24576;
  const int lead = (threadIdx.x & 31) == 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));      // never arrived on: phase parity 1 reads as complete
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + 8), "r"(8));  // 8 warps arrive, nobody waits
  }
  __syncthreads();
  int u = ub;  // uniform value (kernel parameter arithmetic only)
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 32; i++) {
      x[i & 15] = fma(x[i & 15], a, b);
      if ((i + 1) % RATIO == 0) {
        const int k = ((i + 1) / RATIO) & 3;
        if (X == X_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[k]) : "r"(ia), "r"(it));
        if (X == X_IADD) y[k] += y[(k + 1) & 3];
        if (X == X_LOP) y[k] = (y[k] | y[(k + 1) & 3]) ^ y[(k + 2) & 3];
        if (X == X_WAIT)
          asm volatile("{\n.reg .pred P1;\nIB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra IB_DONE;\nbra IB_WAIT;\nIB_DONE:\n}" ::"r"(bar), "r"(1u) : "memory");
        if (X == X_LDS128) asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l0[k]), "=d"(l1[k]) : "r"(saddr + k * 4096));
        if (X == X_LDS64) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(l0[k]) : "r"(saddr + k * 4096));
        if (X == X_LDS128U) asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l0[k]), "=d"(l1[k]) : "r"(saddr + (unsigned)(u & 0x3ff0)));
        if (X == X_ARRIVE) asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n@q mbarrier.arrive.shared::cta.b64 _, [%0];\n}" ::"r"(bar + 8), "r"(lead) : "memory");
        if (X == X_SYNCWARP) __syncwarp();
        if (X == X_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"((float)a), "f"((float)b));
        if (X == X_SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(y[k]));
        if (X == X_UADD) u = u * 3 + ub;
        if (X == X_DADD) asm volatile("add.f64 %0, %0, %1;" : "+d"(l0[k]) : "d"(b));
      }
    }
    if (X == X_LDS128U) u += 16;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s += y[i] + f[i] + l0[i] + l1[i];
  s += u;
  if (s == 12345.678) out[0] = s;
}

template <typename L> static float time_ms(L launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 2; i++) launch();
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

static int g_sms;
static double *g_out;
static bool g_first = true;
template <int X, int RATIO> static void run() {
  const int iters = 4000;
  CK(cudaFuncSetAttribute(issue_kernel<X, RATIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  const float t = time_ms([&] { issue_kernel<X, RATIO><<<g_sms, 256, 32768>>>(g_out, iters, 1.0000001, 1e-9, 3, 16); }, 3);
  const double tf = (double)g_sms * 256 * iters * 32.0 * 2 / t / 1e9;
  printf("%s\"%s_per%d\": %.2f", g_first ? "" : ", ", kNames[X], RATIO, tf);
  g_first = false;
}
template <int X> static void run_all() {
  run<X, 16>();
  run<X, 4>();
  run<X, 2>();
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  CK(cudaMalloc(&g_out, 64));
  printf("{");
  run<X_NONE, 32>();
  run_all<X_IMAD>(); run_all<X_IADD>(); run_all<X_LOP>(); run_all<X_WAIT>(); run_all<X_LDS128>(); run_all<X_LDS64>();
  run_all<X_LDS128U>(); run_all<X_ARRIVE>(); run_all<X_SYNCWARP>(); run_all<X_FFMA>(); run_all<X_SHFL>(); run_all<X_UADD>();
  run_all<X_DADD>();
  printf("}\n");
  return 0;
}
