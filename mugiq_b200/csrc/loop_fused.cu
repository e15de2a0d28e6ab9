// loop_fused.cu — stages 1+2 fused: the eigenvector / displacement loop nest of
// Loop_Mugiq::computeCoarseLoop (/root/reference/lib/loop_mugiq.cpp:455-509).
//
// Schedule "batched" (this file, v0): eigenvectors are processed in batches; per displacement entry the
// whole batch is displaced hop by hop into a ping-pong workspace (one link load serves the batch) and
// contracted with register accumulation over the batch, so the loop buffer is read-modified-written once
// per batch and hop instead of once per eigenvector and hop.  The reference's per-hop blas::zero and the
// two full field copies of Displace::swapAuxDispVec (lib/displace.cpp:47-52,59) are gone: the ping-pong
// buffers swap by pointer.
#include "kernels.cuh"

namespace mugiq_b200 {

constexpr int kLoopBatch = 16;  // eigenvectors displaced together (workspace = 2 * kLoopBatch fields)

static size_t field_bytes(const LatGeom &g, int precision) {
  return (size_t)g.volume * kSpinorLen * 2 * prec_bytes(precision);
}

long long loop_workspace_bytes(const LatGeom &g, int precision, int nvec, const mugiq_b200_disp_entry_t *entries,
                               int nentries) {
  (void)entries;
  if (nentries <= 0) return 0;
  const int nb = nvec < kLoopBatch ? nvec : kLoopBatch;
  return (long long)(2 * (size_t)nb * field_bytes(g, precision));
}

int loop_accumulate(void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec, const void *gauge_d,
                    const mugiq_b200_disp_entry_t *entries, int nentries, int accumulate, void *workspace_d,
                    const LatGeom &g, int precision, cudaStream_t stream) {
  const size_t loop_bytes = (size_t)16 * g.volume * 2 * prec_bytes(precision);  // one loop (16 gammas)
  char *pos = static_cast<char *>(dataPos_d);

  // iL = 0: ultra-local, vR = vL   (lib/loop_mugiq.cpp:499-503)
  int rc = contract_batch(pos, evec_d, nullptr, sigma_h, nvec, accumulate, g, precision, stream);
  if (rc) return rc;
  if (nentries == 0) return MUGIQ_B200_OK;
  if (!workspace_d) return set_error(MUGIQ_B200_EINVAL, "loop_accumulate: workspace_d is NULL");

  const size_t fb = field_bytes(g, precision);
  int iL = 1;
  for (int e = 0; e < nentries; e++) {
    const mugiq_b200_disp_entry_t &en = entries[e];
    const int nL = en.stop - en.start + 1;  // nLoopPerEntry, include/loop_mugiq.h:241
    // the reference zeroes the entry's slots (lib/loop_mugiq.cpp:476) and fills them in the order the
    // contractions happen (dispCount), so slots that no hop reaches (start < 1) stay zero
    if (!accumulate) MUGIQ_CUDA_CHECK(cudaMemsetAsync(pos + (size_t)iL * loop_bytes, 0, (size_t)nL * loop_bytes, stream));
    for (int b0 = 0; b0 < nvec; b0 += kLoopBatch) {
      const int nb = (nvec - b0 < kLoopBatch) ? nvec - b0 : kLoopBatch;
      const void *src[kLoopBatch];
      void *dst[kLoopBatch];
      for (int i = 0; i < nb; i++) src[i] = evec_d[b0 + i];
      int dispCount = 0;
      for (int k = 1; k <= en.stop; k++) {
        for (int i = 0; i < nb; i++) dst[i] = static_cast<char *>(workspace_d) + ((size_t)(k & 1) * nb + i) * fb;
        rc = displace_batch(dst, src, nb, gauge_d, en.dir, en.sign, g, precision, stream);
        if (rc) return rc;
        for (int i = 0; i < nb; i++) src[i] = dst[i];
        if (k >= en.start && dispCount < nL) {
          rc = contract_batch(pos + (size_t)(iL + dispCount) * loop_bytes, evec_d + b0, src, sigma_h + b0, nb, 1, g,
                              precision, stream);
          if (rc) return rc;
          dispCount++;
        }
      }
    }
    iL += nL;
  }
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200
