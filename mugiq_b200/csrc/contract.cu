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
//
// Two kernels: contract_tile_kernel streams 64-site tiles of every eigenvector (pair) of the batch through a
// ring of shared-memory stages filled by bulk TMA (HBM-bound: S or 2S bytes and 192 DFMA per eigvec*site; the
// per-thread strided 128-bit global loads of contract_site_kernel reached 70 % of the HBM peak);
// contract_site_kernel serves fields that are not 16-byte aligned.
#include "kernels.cuh"
#include "tma.cuh"

namespace mugiq_b200 {

constexpr int kCtrTile = 64;    // sites per CTA = threads per CTA
constexpr int kCtrStages = 4;   // shared-memory stages (eigenvectors in flight per CTA: kCtrStages - 1)

// out[(b+K)&3][(a+K)&3] = in[b][a]
template <typename F, int K> __device__ __forceinline__ void unrotate_both(Cplx<F> M[4][4]) {
  Cplx<F> T[4][4];
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int a = 0; a < 4; a++) T[(b + K) & 3][(a + K) & 3] = M[b][a];
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int a = 0; a < 4; a++) M[b][a] = T[b][a];
}

// Thread = site.  The 192-byte site stride would put the 8 lanes of a quarter-warp on two 16-byte bank groups;
// instead each lane reads its spins rotated by k = (lane/2) mod 4 (slot b holds spin (b+k) mod 4), which makes the
// 128-bit shared reads conflict-free and only relabels M; the rotation is undone once after the eigenvector sum.
template <typename F, bool kSame>
__global__ void __launch_bounds__(kCtrTile)
contract_tile_kernel(F *__restrict__ loop, const VecBatch batch, const int accumulate, const LatGeom g, const int nstages) {
  extern __shared__ __align__(128) char smem[];
  constexpr int kS = kSpinorLen * 2 * (int)sizeof(F);
  constexpr int kC = 2 * (int)sizeof(F);
  constexpr int kStage = (kSame ? 1 : 2) * kCtrTile * kS;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + nstages * kStage);  // nstages = min(kCtrStages, nvec)

  const int site0 = blockIdx.x * kCtrTile;
  const int nsites = min(kCtrTile, g.volume - site0);
  const uint32_t bytes = (uint32_t)(nsites * kS);
  const bool active = (int)threadIdx.x < nsites;
  const int x_eo = site0 + threadIdx.x;

  auto issue = [&](int n, int slot) {  // thread 0 only
    char *dst = smem + slot * kStage;
    tma::mbar_expect_tx(&full[slot], kSame ? bytes : 2 * bytes);
    tma::bulk_g2s(dst, static_cast<const char *>(batch.vL[n]) + (size_t)site0 * kS, bytes, &full[slot]);
    if (!kSame) tma::bulk_g2s(dst + kCtrTile * kS, static_cast<const char *>(batch.vR[n]) + (size_t)site0 * kS, bytes, &full[slot]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; s++) tma::mbar_init(&full[s], 1);
    tma::mbar_init_fence();
    tma::fence_async_smem();
    for (int n = 0; n < nstages - 1 || n == 0; n++) issue(n, n);  // nstages <= nvec
  }
  // accumulating call (performLoopContraction is one): fetch the old loop values while the first stage is in flight
  Cplx<F> old[16];
  if (accumulate && active) {
#pragma unroll
    for (int G = 0; G < 16; G++) old[G] = ldg_c<F>(loop + 2 * ((size_t)x_eo + (size_t)g.volume * G));
  }
  __syncthreads();

  Cplx<F> M[4][4];
#pragma unroll
  for (int be = 0; be < 4; be++)
#pragma unroll
    for (int al = 0; al < 4; al++) M[be][al] = make_c<F>(0, 0);

  const int k = (threadIdx.x >> 1) & 3;
  int off[4];
#pragma unroll
  for (int b = 0; b < 4; b++) off[b] = threadIdx.x * kS + ((b + k) & 3) * 3 * kC;

  int slot = 0, pslot = nstages - 1;  // slot of eigenvector n, slot of eigenvector n + nstages - 1
  uint32_t par = 0;
  for (int n = 0; n < batch.nvec; n++) {
    // the slot of eigenvector n+S-1 was read in iteration n-1, before that iteration's barrier
    if (threadIdx.x == 0 && nstages > 1 && n + nstages - 1 < batch.nvec) issue(n + nstages - 1, pslot);
    tma::mbar_wait(&full[slot], par);
    const char *st = smem + slot * kStage;
    const F inv_sigma = (F)batch.inv_sigma[n];
    Cplx<F> l[kSpinorLen], r[kSpinorLen];
    if (active) {
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          l[b * 3 + c] = tma::lds_c<F>(st + off[b] + c * kC);
          if (!kSame) r[b * 3 + c] = tma::lds_c<F>(st + kCtrTile * kS + off[b] + c * kC);
        }
    } else {
#pragma unroll
      for (int i = 0; i < kSpinorLen; i++) l[i] = r[i] = make_c<F>(0, 0);
    }
    __syncthreads();  // the stage may be refilled
    if (kSame) {
#pragma unroll
      for (int i = 0; i < kSpinorLen; i++) r[i] = l[i];
    }
#pragma unroll
    for (int i = 0; i < kSpinorLen; i++) {
      l[i].re *= inv_sigma;
      l[i].im *= inv_sigma;
    }
#pragma unroll
    for (int be = 0; be < 4; be++)
#pragma unroll
      for (int al = 0; al < 4; al++)
#pragma unroll
        for (int c = 0; c < 3; c++) cmac_conj(M[be][al], l[be * 3 + c], r[al * 3 + c]);
    if (++slot == nstages) {
      slot = 0;
      par ^= 1u;
    }
    if (++pslot == nstages) pslot = 0;
  }
  if (!active) return;
  if (k == 1) unrotate_both<F, 1>(M);
  if (k == 2) unrotate_both<F, 2>(M);
  if (k == 3) unrotate_both<F, 3>(M);

  Cplx<F> T[16];
  gamma_project(T, M);
#pragma unroll
  for (int G = 0; G < 16; G++) {
    F *p = loop + 2 * ((size_t)x_eo + (size_t)g.volume * G);
    Cplx<F> out = T[G];
    if (accumulate) {
      out.re += old[G].re;
      out.im += old[G].im;
    }
    st_c<F>(p, out);
  }
}

template <typename F, bool kSame>
__global__ void __launch_bounds__(128)
contract_site_kernel(F *__restrict__ loop, const VecBatch batch, const int accumulate, const LatGeom g) {
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

template <typename F, bool kSame>
static int launch_tile(void *loop_d, const VecBatch &batch, int accumulate, const LatGeom &g, cudaStream_t stream) {
  const int nstages = batch.nvec < kCtrStages ? batch.nvec : kCtrStages;  // a single pair leaves room for more CTAs per SM
  const int smem = nstages * (kSame ? 1 : 2) * kCtrTile * kSpinorLen * 2 * (int)sizeof(F) + kCtrStages * 8;
  MUGIQ_CUDA_CHECK(cudaFuncSetAttribute(contract_tile_kernel<F, kSame>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int blocks = (g.volume + kCtrTile - 1) / kCtrTile;
  contract_tile_kernel<F, kSame><<<blocks, kCtrTile, smem, stream>>>((F *)loop_d, batch, accumulate, g, nstages);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

template <typename F>
static int launch_contract(void *loop_d, const VecBatch &batch, bool same, int accumulate, const LatGeom &g,
                           cudaStream_t stream) {
  const int threads = 128;
  const int blocks = (g.volume + threads - 1) / threads;
  // algorithmic bytes: S (ultra-local) or 2S per eigvec·site, accumulator written once (+ read if accumulating)
  const double S = kSpinorLen * 2.0 * sizeof(F), A = 16 * 2.0 * sizeof(F);
  ProfScope prof(K_CONTRACT, stream, (double)g.volume * (batch.nvec * (same ? S : 2 * S) + (accumulate ? 2 * A : A)),
                 (double)g.volume * batch.nvec * 2.0 * (192 + 24));
  bool aligned = true;  // bulk copies need 16-byte aligned fields
  for (int i = 0; i < batch.nvec; i++)
    if (((uintptr_t)batch.vL[i] | (uintptr_t)batch.vR[i]) & 15) aligned = false;
  if (aligned) return same ? launch_tile<F, true>(loop_d, batch, accumulate, g, stream)
                           : launch_tile<F, false>(loop_d, batch, accumulate, g, stream);
  if (same)
    contract_site_kernel<F, true><<<blocks, threads, 0, stream>>>((F *)loop_d, batch, accumulate, g);
  else
    contract_site_kernel<F, false><<<blocks, threads, 0, stream>>>((F *)loop_d, batch, accumulate, g);
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
