// momproj.cu — stage 4: Fourier momentum projection as a complex GEMM against the phase matrix.
//
//   dataMom(M x N) = dataPosMP(M x K) * phase(K x N),  column-major, M = Lt*16*nLoop, N = Nmom, K = V3.
// Replaces cublasZgemm / cublasCgemm in /root/reference/lib/loop_mugiq.cpp:358-377.
//
// The problem is tall and skinny (N = 1..33 momenta): the M x K operand is streamed from HBM exactly
// once and never reused, so it goes global -> registers directly in MMA-fragment order (no shared-memory
// staging to pay for); the small K x N phase operand is re-read through L1/L2.
//
// FP64 path: mma.sync.m8n8k4 DMMA tiles (tcgen05 has no FP64 kind; mma.sync is the FP64 tensor route on
// sm_100).  The complex product is mapped onto real tiles as C^T = P_emb^T * A^T:
//   tile rows  r = (momentum n0 + r/2, component r%2)   (4 momenta x {re,im})
//   tile cols  8 consecutive rows m of dataPosMP
//   k          4 consecutive spatial sites, taken twice: once with Re(A) and once with Im(A)
//   pass 1:  a = (comp ? Im P : Re P),  b = Re A        pass 2:  a = (comp ? Re P : -Im P),  b = Im A
// so every lane loads exactly one 128-bit complex of A and one of P per k-step and feeds both passes.
// FP32 path: SIMT kernel (the headline configuration is FP64).
// Split-K partial sums go to a workspace and are reduced in a fixed order (deterministic result).
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"

namespace mugiq_b200 {

constexpr int kMT = 4;        // 8-row m-tiles per warp
constexpr int kWarps = 4;     // warps per CTA (consecutive m ranges, same k range)
constexpr int kRowsPerCta = kWarps * kMT * 8;

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// NT = number of 4-momentum n-tiles handled by one warp
template <int NT>
__global__ void __launch_bounds__(kWarps * 32)
momproj_dmma_kernel(double *__restrict__ partial, const double *__restrict__ A, const double *__restrict__ P,
                    const long long M, const int N, const long long K, const long long kchunk) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gi = lane >> 2;  // row of the a-fragment / column of the b-fragment
  const int j = lane & 3;    // k index inside the step
  const long long m0 = (long long)blockIdx.x * kRowsPerCta + warp * (kMT * 8);
  const int n0 = blockIdx.z * (NT * 4);
  const long long kbeg = (long long)blockIdx.y * kchunk;
  const long long kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
  const int comp = gi & 1;

  double acc[kMT][NT][2];
#pragma unroll
  for (int i = 0; i < kMT; i++)
#pragma unroll
    for (int t = 0; t < NT; t++) acc[i][t][0] = acc[i][t][1] = 0.0;

  for (long long k0 = kbeg; k0 < kend; k0 += 4) {
    const long long k = k0 + j;
    const bool kok = k < kend;
    double2 d[kMT];
#pragma unroll
    for (int i = 0; i < kMT; i++) {
      const long long m = m0 + i * 8 + gi;
      d[i] = (kok && m < M) ? __ldg(reinterpret_cast<const double2 *>(A) + (m + M * k)) : make_double2(0.0, 0.0);
    }
    double a1[NT], a2[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const int n = n0 + t * 4 + (gi >> 1);
      const double2 p =
          (kok && n < N) ? __ldg(reinterpret_cast<const double2 *>(P) + (k + K * n)) : make_double2(0.0, 0.0);
      a1[t] = comp ? p.y : p.x;
      a2[t] = comp ? p.x : -p.y;
    }
#pragma unroll
    for (int i = 0; i < kMT; i++)
#pragma unroll
      for (int t = 0; t < NT; t++) {
        dmma_m8n8k4(acc[i][t][0], acc[i][t][1], a1[t], d[i].x);
        dmma_m8n8k4(acc[i][t][0], acc[i][t][1], a2[t], d[i].y);
      }
  }

  // c-fragment: row gi -> (n, comp), columns 2j, 2j+1 -> m
  double *out = partial + 2 * (size_t)blockIdx.y * (size_t)M * N;
#pragma unroll
  for (int i = 0; i < kMT; i++)
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const int n = n0 + t * 4 + (gi >> 1);
      if (n < N) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const long long m = m0 + i * 8 + 2 * j + e;
          if (m < M) out[2 * (m + M * n) + comp] = acc[i][t][e];
        }
      }
    }
}

// SIMT version (FP32, and FP64 comparator selected with MUGIQ_B200_MOMPROJ=simt)
constexpr int kSimtN = 4;
template <typename F>
__global__ void __launch_bounds__(128)
momproj_simt_kernel(F *__restrict__ partial, const F *__restrict__ A, const F *__restrict__ P, const long long M,
                    const int N, const long long K, const long long kchunk) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = blockIdx.z * kSimtN;
  const long long kbeg = (long long)blockIdx.y * kchunk;
  const long long kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
  if (m >= M) return;
  Cplx<F> acc[kSimtN];
#pragma unroll
  for (int t = 0; t < kSimtN; t++) acc[t] = make_c<F>(0, 0);
  for (long long k = kbeg; k < kend; k++) {
    const Cplx<F> a = ldg_c<F>(A + 2 * (m + M * k));
#pragma unroll
    for (int t = 0; t < kSimtN; t++)
      if (n0 + t < N) cmac(acc[t], a, ldg_c<F>(P + 2 * (k + K * (n0 + t))));
  }
  F *out = partial + 2 * (size_t)blockIdx.y * (size_t)M * N;
#pragma unroll
  for (int t = 0; t < kSimtN; t++)
    if (n0 + t < N) st_c<F>(out + 2 * (m + M * (n0 + t)), acc[t]);
}

template <typename F>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(F *__restrict__ out, const F *__restrict__ partial, const long long nreal, const int ksplit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nreal) return;
  F s = 0;
  for (int ks = 0; ks < ksplit; ks++) s += partial[(size_t)ks * nreal + i];
  out[i] = s;
}

static int pick_ksplit(long long M, int N, long long K, int rows_per_cta, int ncols_per_cta) {
  const long long ctas_mn = ((M + rows_per_cta - 1) / rows_per_cta) * ((N + ncols_per_cta - 1) / ncols_per_cta);
  long long want = (148LL * 8 + ctas_mn - 1) / ctas_mn;  // ~8 CTAs per SM in flight
  long long maxk = K / 64;                                // at least 64 k per chunk
  if (maxk < 1) maxk = 1;
  if (want > maxk) want = maxk;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return (int)want;
}

static bool use_simt() {
  const char *e = getenv("MUGIQ_B200_MOMPROJ");
  return e && strcmp(e, "simt") == 0;
}

static int nt_for(int N) {
  const int ntiles = (N + 3) / 4;  // 4 momenta per n-tile; more than 9 tiles -> several passes over A
  return ntiles >= 9 ? 9 : ntiles;
}

static void ksplit_for(long long M, int N, long long K, int precision, int *ksplit, long long *kchunk) {
  int ks;
  if (precision == MUGIQ_B200_PREC_DOUBLE && !use_simt())
    ks = pick_ksplit(M, N, K, kRowsPerCta, nt_for(N) * 4);
  else
    ks = pick_ksplit(M, N, K, 128, kSimtN);
  long long kc = (K + ks - 1) / ks;
  kc = (kc + 3) / 4 * 4;
  ks = (int)((K + kc - 1) / kc);
  *ksplit = ks;
  *kchunk = kc;
}

long long momproj_workspace_bytes(long long M, int N, long long K, int precision) {
  int ks;
  long long kc;
  ksplit_for(M, N, K, precision, &ks, &kc);
  return (long long)ks * M * N * 2 * (long long)prec_bytes(precision);
}

int momproj(void *mom_d, const void *posMP_d, const void *phase_d, long long M, int N, long long K, int precision,
            void *workspace_d, cudaStream_t stream) {
  int ks;
  long long kc;
  ksplit_for(M, N, K, precision, &ks, &kc);
  if (!workspace_d) return set_error(MUGIQ_B200_EINVAL, "momproj: workspace_d is NULL");
  {
  // algorithmic bytes: A once, phase once, result once; flops 8*M*N*K are reported by the caller
  ProfScope prof(K_MOMPROJ, stream, 2.0 * prec_bytes(precision) * ((double)M * K + (double)K * N + (double)M * N),
                 8.0 * (double)M * N * (double)K);
  if (precision == MUGIQ_B200_PREC_DOUBLE && !use_simt()) {
    const int NT = nt_for(N);
    const dim3 grid((unsigned)((M + kRowsPerCta - 1) / kRowsPerCta), ks, (N + NT * 4 - 1) / (NT * 4));
    const dim3 block(kWarps * 32);
    double *ws = (double *)workspace_d;
    const double *A = (const double *)posMP_d, *P = (const double *)phase_d;
    switch (NT) {
#define MUGIQ_NT_CASE(nt) \
  case nt: momproj_dmma_kernel<nt><<<grid, block, 0, stream>>>(ws, A, P, M, N, K, kc); break;
      MUGIQ_NT_CASE(9) MUGIQ_NT_CASE(8) MUGIQ_NT_CASE(7) MUGIQ_NT_CASE(6) MUGIQ_NT_CASE(5)
      MUGIQ_NT_CASE(4) MUGIQ_NT_CASE(3) MUGIQ_NT_CASE(2)
      default: momproj_dmma_kernel<1><<<grid, block, 0, stream>>>(ws, A, P, M, N, K, kc); break;
#undef MUGIQ_NT_CASE
    }
    MUGIQ_LAUNCH_CHECK();
  } else {
    const dim3 grid((unsigned)((M + 127) / 128), ks, (N + kSimtN - 1) / kSimtN);
    if (precision == MUGIQ_B200_PREC_DOUBLE)
      momproj_simt_kernel<double><<<grid, 128, 0, stream>>>((double *)workspace_d, (const double *)posMP_d,
                                                             (const double *)phase_d, M, N, K, kc);
    else
      momproj_simt_kernel<float><<<grid, 128, 0, stream>>>((float *)workspace_d, (const float *)posMP_d,
                                                            (const float *)phase_d, M, N, K, kc);
    MUGIQ_LAUNCH_CHECK();
  }
  }
  const long long nreal = 2 * M * N;
  const int blocks = (int)((nreal + 255) / 256);
  ProfScope prof2(K_SPLITK_REDUCE, stream, (double)nreal * (ks + 1) * prec_bytes(precision));
  if (precision == MUGIQ_B200_PREC_DOUBLE)
    splitk_reduce_kernel<double><<<blocks, 256, 0, stream>>>((double *)mom_d, (const double *)workspace_d, nreal, ks);
  else
    splitk_reduce_kernel<float><<<blocks, 256, 0, stream>>>((float *)mom_d, (const float *)workspace_d, nreal, ks);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200
