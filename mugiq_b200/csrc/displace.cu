// displace.cu — stage 2: covariant displacement of colour-spinor fields by SU(3) links.
//
// Replaces /root/reference/lib/mugiq_displace_kernels.cu:156-185 (covariantDisplacementVector_kernel) and
// its wrapper lib/contract_wrappers.cu:171-198:
//   plus : dst(x) = U_mu(x)            * src(x+mu)      (lib/mugiq_displace_kernels.cu:58-59,124-129)
//   minus: dst(x) = U_mu(x-mu)^dagger  * src(x-mu)      (lib/mugiq_displace_kernels.cu:60-63,137-147)
// with periodic wrap (single-process lattice: the reference's extended-halo field has border 0,
// lib/displace.cpp:16) and Link*Vector = sum_c' U(c,c') v(s,c').
// Several eigenvectors are displaced per launch so that one link load serves the whole batch.
#include "kernels.cuh"

namespace mugiq_b200 {

constexpr int kDispBatch = 8;  // eigenvectors displaced per thread with one link load

struct DispBatch {
  const void *src[kDispBatch];
  void *dst[kDispBatch];
  int nvec;
};

template <typename F>
__global__ void __launch_bounds__(128)
displace_kernel(const DispBatch batch, const F *__restrict__ gauge, const int dir, const int sign, const LatGeom g) {
  const int x_eo = blockIdx.x * blockDim.x + threadIdx.x;
  if (x_eo >= g.volume) return;
  const int pty = x_eo >= g.volumeCB;
  const int x_cb = x_eo - pty * g.volumeCB;

  int x[4];
  get_coords(x, x_cb, pty, g);
  // neighbour site x +- mu (periodic), always of the opposite parity
  x[dir] = sign ? (x[dir] + 1 == g.L[dir] ? 0 : x[dir] + 1) : (x[dir] == 0 ? g.L[dir] - 1 : x[dir] - 1);
  const int nbr_cb = cb_index(x, g);
  const int nbr_eo = nbr_cb + (1 - pty) * g.volumeCB;

  // link: U_mu(x) for plus, U_mu(x-mu) (to be daggered) for minus
  const size_t link_site = sign ? (size_t)x_eo : (size_t)nbr_eo;
  const F *pu = gauge + 2 * kLinkLen * ((size_t)dir * g.volume + link_site);
  Cplx<F> U[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const Cplx<F> u = ldg_c<F>(pu + 2 * (r * 3 + c));
      if (sign) {
        U[r][c] = u;
      } else {  // Hermitian conjugate
        U[c][r] = make_c<F>(u.re, -u.im);
      }
    }

  for (int n = 0; n < batch.nvec; n++) {
    const F *ps = static_cast<const F *>(batch.src[n]) + (size_t)nbr_eo * (2 * kSpinorLen);
    F *pd = static_cast<F *>(batch.dst[n]) + (size_t)x_eo * (2 * kSpinorLen);
    Cplx<F> v[kSpinorLen];
#pragma unroll
    for (int k = 0; k < kSpinorLen; k++) v[k] = ldg_c<F>(ps + 2 * k);
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        Cplx<F> acc = make_c<F>(0, 0);
#pragma unroll
        for (int cp = 0; cp < 3; cp++) cmac(acc, U[c][cp], v[s * 3 + cp]);
        st_c<F>(pd + 2 * (s * 3 + c), acc);
      }
  }
}

int displace_batch(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir, int sign,
                   const LatGeom &g, int precision, cudaStream_t stream) {
  const int threads = 128;
  const int blocks = (g.volume + threads - 1) / threads;
  for (int done = 0; done < nvec; done += kDispBatch) {
    DispBatch batch;
    batch.nvec = (nvec - done < kDispBatch) ? nvec - done : kDispBatch;
    for (int i = 0; i < batch.nvec; i++) {
      batch.src[i] = src_d[done + i];
      batch.dst[i] = dst_d[done + i];
    }
    const double pb = (double)prec_bytes(precision);
    ProfScope prof(K_DISPLACE, stream, (double)g.volume * pb * 2.0 * (batch.nvec * 2.0 * kSpinorLen + kLinkLen));
    if (precision == MUGIQ_B200_PREC_DOUBLE)
      displace_kernel<double><<<blocks, threads, 0, stream>>>(batch, (const double *)gauge_d, dir, sign, g);
    else
      displace_kernel<float><<<blocks, threads, 0, stream>>>(batch, (const float *)gauge_d, dir, sign, g);
    MUGIQ_LAUNCH_CHECK();
  }
  return MUGIQ_B200_OK;
}

int displace(void *dst_d, const void *src_d, const void *gauge_d, int dir, int sign, const LatGeom &g,
             int precision, cudaStream_t stream) {
  return displace_batch(&dst_d, &src_d, 1, gauge_d, dir, sign, g, precision, stream);
}

}  // namespace mugiq_b200
