// contract.cu — stage 1: 16-gamma loop contraction (per-pair and batched entry points).
//
// Replaces /root/reference/lib/mugiq_contract_kernels.cu:45-122 (loopContract_kernel) and its wrapper
// lib/contract_wrappers.cu:88-115.  Differences in structure (not in result):
//  * one thread owns one site and all 16 (be,al) spin pairs; the reference spends 16 z-threads per site
//    of which one loads, and stages everything through shared memory;
//  * the eigenvector sum runs inside the kernel with the 4x4 colour-traced spin matrix kept in
//    registers, so the loop buffer is written once per launch instead of read-modified-written once
//    per eigenvector (lib/mugiq_contract_kernels.cu:120);
//  * the gamma projection is applied once, after the eigenvector sum (it is linear), using only
//    adds/swaps because every coefficient is +-1 or +-i (include/gamma.h:33-48).
#include "kernels.cuh"

namespace mugiq_b200 {

template <typename F, bool kSame>
__global__ void __launch_bounds__(128)
contract_batch_kernel(F *__restrict__ loop, const VecBatch batch, const int accumulate, const LatGeom g) {
  const int x_eo = blockIdx.x * blockDim.x + threadIdx.x;  // full-site index, parity-major
  if (x_eo >= g.volume) return;

  Cplx<F> M[4][4];
#pragma unroll
  for (int be = 0; be < 4; be++)
#pragma unroll
    for (int al = 0; al < 4; al++) M[be][al] = make_c<F>(0, 0);

  for (int n = 0; n < batch.nvec; n++) {
    const F *pl = static_cast<const F *>(batch.vL[n]) + (size_t)x_eo * (2 * kSpinorLen);
    const F inv_sigma = (F)batch.inv_sigma[n];
    Cplx<F> l[kSpinorLen], r[kSpinorLen];
#pragma unroll
    for (int k = 0; k < kSpinorLen; k++) l[k] = ldg_c<F>(pl + 2 * k);
    if (kSame) {
#pragma unroll
      for (int k = 0; k < kSpinorLen; k++) r[k] = l[k];
    } else {
      const F *pr = static_cast<const F *>(batch.vR[n]) + (size_t)x_eo * (2 * kSpinorLen);
#pragma unroll
      for (int k = 0; k < kSpinorLen; k++) r[k] = ldg_c<F>(pr + 2 * k);
    }
#pragma unroll
    for (int k = 0; k < kSpinorLen; k++) {
      l[k].re *= inv_sigma;
      l[k].im *= inv_sigma;
    }
    // M[be][al] += sum_c conj(vL[be,c]) vR[al,c]     (lib/mugiq_contract_kernels.cu:103-105)
#pragma unroll
    for (int be = 0; be < 4; be++)
#pragma unroll
      for (int al = 0; al < 4; al++)
#pragma unroll
        for (int c = 0; c < 3; c++) cmac_conj(M[be][al], l[be * 3 + c], r[al * 3 + c]);
  }

  Cplx<F> T[16];
  gamma_project(T, M);
#pragma unroll
  for (int G = 0; G < 16; G++) {
    F *p = loop + 2 * ((size_t)x_eo + (size_t)g.volume * G);
    Cplx<F> out = T[G];
    if (accumulate) {
      const Cplx<F> old = ldg_c<F>(p);
      out.re += old.re;
      out.im += old.im;
    }
    st_c<F>(p, out);
  }
}

template <typename F>
static int launch_contract(void *loop_d, const VecBatch &batch, bool same, int accumulate, const LatGeom &g,
                           cudaStream_t stream) {
  const int threads = 128;
  const int blocks = (g.volume + threads - 1) / threads;
  // algorithmic bytes: S (ultra-local) or 2S per eigvec·site, accumulator written once (+ read if accumulating)
  const double S = kSpinorLen * 2.0 * sizeof(F), A = 16 * 2.0 * sizeof(F);
  ProfScope prof(K_CONTRACT, stream, (double)g.volume * (batch.nvec * (same ? S : 2 * S) + (accumulate ? 2 * A : A)));
  if (same)
    contract_batch_kernel<F, true><<<blocks, threads, 0, stream>>>((F *)loop_d, batch, accumulate, g);
  else
    contract_batch_kernel<F, false><<<blocks, threads, 0, stream>>>((F *)loop_d, batch, accumulate, g);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

int contract_batch(void *loop_d, const void *const *vL, const void *const *vR, const double *sigma, int nvec,
                   int accumulate, const LatGeom &g, int precision, cudaStream_t stream) {
  int done = 0;
  while (done < nvec) {
    VecBatch batch;
    batch.nvec = (nvec - done < kMaxBatch) ? nvec - done : kMaxBatch;
    for (int i = 0; i < batch.nvec; i++) {
      batch.vL[i] = vL[done + i];
      batch.vR[i] = vR ? vR[done + i] : vL[done + i];
      batch.inv_sigma[i] = inv_sigma_of(sigma[done + i], precision);
    }
    const int acc = accumulate || done > 0;
    int rc = (precision == MUGIQ_B200_PREC_DOUBLE)
                 ? launch_contract<double>(loop_d, batch, vR == nullptr, acc, g, stream)
                 : launch_contract<float>(loop_d, batch, vR == nullptr, acc, g, stream);
    if (rc) return rc;
    done += batch.nvec;
  }
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200
