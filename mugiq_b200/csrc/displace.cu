// displace.cu — stage 2: covariant displacement of colour-spinor fields by SU(3) links.
//
// Replaces /root/reference/lib/mugiq_displace_kernels.cu:156-185 (covariantDisplacementVector_kernel) and
// its wrapper lib/contract_wrappers.cu:171-198:
//   plus : dst(x) = U_mu(x)            * src(x+mu)      (lib/mugiq_displace_kernels.cu:58-59,124-129)
//   minus: dst(x) = U_mu(x-mu)^dagger  * src(x-mu)      (lib/mugiq_displace_kernels.cu:60-63,137-147)
// with periodic wrap (single-process lattice: the reference's extended-halo field has border 0,
// lib/displace.cpp:16) and Link*Vector = sum_c' U(c,c') v(s,c').
//
// The kernel is pure HBM streaming (2S + U = 528 B and 144 DFMA per site), so it is built around whole-tile
// bulk-TMA copies instead of per-thread strided 128-bit accesses (which reached 2.4 TB/s):
//  * CTA tile = R consecutive-y half-rows of one parity at fixed (z,t): in the even/odd site-major layout the
//    tile, its neighbour tile (opposite parity, rows shifted by +-1 in the displacement direction; the SAME rows
//    for an x displacement) and their links are each ONE contiguous run in memory (two when the y shift wraps),
//    fetched by cp.async.bulk into shared memory with an mbarrier transaction count;
//  * thread = (site, spin): 3 complex in, 3 complex out, 48-B stride => conflict-free 128-bit shared accesses;
//  * the result tile leaves by a bulk store (shared -> global) while the next eigenvector of the batch is computed;
//  * several eigenvectors are displaced per launch so that one link tile serves the whole batch.
// Lattices whose half-row is not a multiple of 16 bytes (FP32 with odd Lx/2) take the per-site kernel below.
#include "kernels.cuh"
#include "tma.cuh"

namespace mugiq_b200 {

constexpr int kDispBatch = 32;      // eigenvectors displaced per launch with one link load
constexpr int kDispThreads = 256;
constexpr int kDispTileSites = 64;  // target sites per tile (12 KB spinor + 9 KB link copies in FP64)

struct DispBatch {
  const void *src[kDispBatch];
  void *dst[kDispBatch];
  int nvec;
};

struct DispTiling {
  int R;          // half-rows per tile (divides Ly)
  int tilesY;     // Ly / R
  int sites;      // R * Lh
  int ntiles;     // 2 * Lt * Lz * tilesY
};

// ---- per-site kernel: any lattice ------------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(128)
displace_site_kernel(const DispBatch batch, const F *__restrict__ gauge, const int dir, const int sign, const LatGeom g) {
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

// ---- tile kernel: bulk-TMA staging ------------------------------------------------------------------------------
// Shared memory: [link tile][in 0][in 1][out 0][out 1], every block `sites` sites long, then two mbarriers.
template <typename F>
__global__ void __launch_bounds__(kDispThreads)
displace_tile_kernel(const DispBatch batch, const F *__restrict__ gauge, const int dir, const int sign, const LatGeom g,
                     const DispTiling tl) {
  extern __shared__ __align__(128) char smem[];
  constexpr int kS = kSpinorLen * 2 * (int)sizeof(F);  // bytes per spinor site
  constexpr int kU = kLinkLen * 2 * (int)sizeof(F);    // bytes per link
  constexpr int kC = 2 * (int)sizeof(F);               // bytes per complex
  const int sites = tl.sites;
  const int ubytes = (sites * kU + 127) & ~127, sbytes = (sites * kS + 127) & ~127;
  char *u_s = smem;
  char *in_s = u_s + ubytes;
  char *out_s = in_s + 2 * sbytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(out_s + 2 * sbytes);

  // tile -> (parity, t, z, y0)
  int b = blockIdx.x;
  const int perParity = tl.ntiles >> 1;
  const int pty = b >= perParity;
  b -= pty * perParity;
  const int ty = b % tl.tilesY;
  const int zt = b / tl.tilesY;
  const int z = zt % g.L[2], t = zt / g.L[2];
  const int y0 = ty * tl.R;
  const int row0 = y0 + g.L[1] * zt;  // lexicographic row index of the tile's first half-row

  // neighbour rows: run `a` of nrow[a] rows starting at lexicographic row grow[a], landing at tile row srow[a]
  int nrun = 1, srow[2] = {0, 0}, grow[2] = {row0, 0}, nrow[2] = {tl.R, 0};
  const int sh = sign ? 1 : -1;
  if (dir == 1) {
    if (sign && y0 + tl.R == g.L[1]) {  // last tile row wraps to y = 0
      nrow[0] = tl.R - 1;
      grow[0] = row0 + 1;
      srow[1] = tl.R - 1;
      grow[1] = g.L[1] * zt;
      nrow[1] = 1;
      nrun = 2;
    } else if (!sign && y0 == 0) {  // first tile row comes from y = Ly-1
      srow[0] = 0;
      grow[0] = g.L[1] - 1 + g.L[1] * zt;
      nrow[0] = 1;
      srow[1] = 1;
      grow[1] = row0;
      nrow[1] = tl.R - 1;
      nrun = 2;
    } else {
      grow[0] = row0 + sh;
    }
  } else if (dir == 2) {
    const int zn = (z + sh + g.L[2]) % g.L[2];
    grow[0] = y0 + g.L[1] * (zn + g.L[2] * t);
  } else if (dir == 3) {
    const int tn = (t + sh + g.L[3]) % g.L[3];
    grow[0] = y0 + g.L[1] * (z + g.L[2] * tn);
  }
  const size_t own_site0 = (size_t)pty * g.volumeCB + (size_t)row0 * g.Lh;  // x_eo of the tile's first site
  const size_t nbr_par0 = (size_t)(1 - pty) * g.volumeCB;

  auto load_vec = [&](int n, int slot) {  // thread 0 only
    const char *src = static_cast<const char *>(batch.src[n]);
    for (int a = 0; a < nrun; a++)
      if (nrow[a] > 0)
        tma::bulk_g2s(in_s + slot * sbytes + srow[a] * g.Lh * kS, src + (nbr_par0 + (size_t)grow[a] * g.Lh) * kS,
                      (uint32_t)(nrow[a] * g.Lh * kS), &full[slot]);
  };

  if (threadIdx.x == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::mbar_init_fence();
    tma::fence_async_smem();
    const char *gl = reinterpret_cast<const char *>(gauge) + (size_t)dir * g.volume * kU;
    tma::mbar_expect_tx(&full[0], (uint32_t)(sites * (kU + kS)));
    if (sign) {
      tma::bulk_g2s(u_s, gl + own_site0 * kU, (uint32_t)(sites * kU), &full[0]);
    } else {
      for (int a = 0; a < nrun; a++)
        if (nrow[a] > 0)
          tma::bulk_g2s(u_s + srow[a] * g.Lh * kU, gl + (nbr_par0 + (size_t)grow[a] * g.Lh) * kU,
                        (uint32_t)(nrow[a] * g.Lh * kU), &full[0]);
    }
    load_vec(0, 0);
    if (batch.nvec > 1) {
      tma::mbar_expect_tx(&full[1], (uint32_t)(sites * kS));
      load_vec(1, 1);
    }
  }
  __syncthreads();  // barrier initialisation visible to the waiters

  const int items = sites * 4;  // (site, spin)
  const int rowpar = (z + t + pty) & 1;
  for (int n = 0; n < batch.nvec; n++) {
    const int slot = n & 1;
    tma::mbar_wait(&full[slot], (uint32_t)((n >> 1) & 1));
    const char *vin = in_s + slot * sbytes;
    char *vout = out_s + slot * sbytes;
    for (int it = threadIdx.x; it < items; it += kDispThreads) {
      const int j = it >> 2, s = it & 3;
      int jn = j;
      if (dir == 0) {  // x neighbour: same half-row index or the next / previous one, depending on the row's odd bit
        const int a = j / g.Lh, i = j - a * g.Lh;
        const int odd = (y0 + a + rowpar) & 1;
        int in_ = i;
        if (sign) {
          if (odd) in_ = (i + 1 == g.Lh) ? 0 : i + 1;
        } else {
          if (!odd) in_ = (i == 0) ? g.Lh - 1 : i - 1;
        }
        jn = a * g.Lh + in_;
      }
      const char *pu = u_s + (sign ? j : jn) * kU;
      const char *pv = vin + jn * kS + s * 3 * kC;
      const Cplx<F> v0 = tma::lds_c<F>(pv), v1 = tma::lds_c<F>(pv + kC), v2 = tma::lds_c<F>(pv + 2 * kC);
      Cplx<F> r[3];
      if (sign) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
          r[c] = cmul(tma::lds_c<F>(pu + (c * 3 + 0) * kC), v0);
          cmac(r[c], tma::lds_c<F>(pu + (c * 3 + 1) * kC), v1);
          cmac(r[c], tma::lds_c<F>(pu + (c * 3 + 2) * kC), v2);
        }
      } else {  // Hermitian conjugate: (U^dag)[c][cp] = conj(U[cp][c])
#pragma unroll
        for (int c = 0; c < 3; c++) {
          Cplx<F> u = tma::lds_c<F>(pu + (0 * 3 + c) * kC);
          u.im = -u.im;
          r[c] = cmul(u, v0);
          u = tma::lds_c<F>(pu + (1 * 3 + c) * kC);
          u.im = -u.im;
          cmac(r[c], u, v1);
          u = tma::lds_c<F>(pu + (2 * 3 + c) * kC);
          u.im = -u.im;
          cmac(r[c], u, v2);
        }
      }
      char *po = vout + j * kS + s * 3 * kC;
      tma::sts_c<F>(po, r[0]);
      tma::sts_c<F>(po + kC, r[1]);
      tma::sts_c<F>(po + 2 * kC, r[2]);
    }
    tma::fence_async_smem();                        // result tile visible to the bulk store
    if (threadIdx.x == 0) tma::bulk_wait_read_all();  // the previous store (other out buffer) has left shared memory
    __syncthreads();
    if (threadIdx.x == 0) {
      tma::bulk_s2g(static_cast<char *>(batch.dst[n]) + own_site0 * kS, vout, (uint32_t)(sites * kS));
      tma::bulk_commit();
      if (n + 2 < batch.nvec) {  // every thread has finished reading in[slot]
        tma::mbar_expect_tx(&full[slot], (uint32_t)(sites * kS));
        load_vec(n + 2, slot);
      }
    }
  }
  if (threadIdx.x == 0) tma::bulk_wait_all();
}

static bool disp_tiling(DispTiling &tl, const LatGeom &g, int precision) {
  const int pb = (int)prec_bytes(precision);
  if ((g.Lh * kLinkLen * 2 * pb) % 16 != 0 || (g.Lh * kSpinorLen * 2 * pb) % 16 != 0) return false;
  int R = 1;
  for (int r = 1; r <= g.L[1]; r++)
    if (g.L[1] % r == 0 && r * g.Lh <= kDispTileSites) R = r;
  tl.R = R;
  tl.tilesY = g.L[1] / R;
  tl.sites = R * g.Lh;
  tl.ntiles = 2 * g.L[3] * g.L[2] * tl.tilesY;
  return true;
}

static size_t disp_smem_bytes(const DispTiling &tl, int precision) {
  const size_t pb = prec_bytes(precision);
  const size_t ub = ((size_t)tl.sites * kLinkLen * 2 * pb + 127) & ~(size_t)127;
  const size_t sb = ((size_t)tl.sites * kSpinorLen * 2 * pb + 127) & ~(size_t)127;
  return ub + 4 * sb + 16;
}

template <typename F>
static int launch_tile(const DispBatch &batch, const void *gauge_d, int dir, int sign, const LatGeom &g, const DispTiling &tl,
                       size_t smem, cudaStream_t stream) {
  MUGIQ_CUDA_CHECK(cudaFuncSetAttribute(displace_tile_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  displace_tile_kernel<F><<<tl.ntiles, kDispThreads, smem, stream>>>(batch, (const F *)gauge_d, dir, sign, g, tl);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

int displace_batch(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir, int sign,
                   const LatGeom &g, int precision, cudaStream_t stream) {
  DispTiling tl;
  const bool tiled = disp_tiling(tl, g, precision);
  const size_t smem = tiled ? disp_smem_bytes(tl, precision) : 0;
  bool use_tile = tiled && smem <= 200 * 1024 && ((uintptr_t)gauge_d & 15) == 0;
  for (int i = 0; i < nvec; i++)  // bulk copies need 16-byte aligned fields
    if (((uintptr_t)dst_d[i] | (uintptr_t)src_d[i]) & 15) use_tile = false;
  for (int done = 0; done < nvec; done += kDispBatch) {
    DispBatch batch;
    batch.nvec = (nvec - done < kDispBatch) ? nvec - done : kDispBatch;
    for (int i = 0; i < batch.nvec; i++) {
      batch.src[i] = src_d[done + i];
      batch.dst[i] = dst_d[done + i];
    }
    const double pb = (double)prec_bytes(precision);
    ProfScope prof(K_DISPLACE, stream, (double)g.volume * pb * 2.0 * (batch.nvec * 2.0 * kSpinorLen + kLinkLen),
                   (double)g.volume * batch.nvec * 288.0);
    if (use_tile) {
      const int rc = (precision == MUGIQ_B200_PREC_DOUBLE)
                         ? launch_tile<double>(batch, gauge_d, dir, sign, g, tl, smem, stream)
                         : launch_tile<float>(batch, gauge_d, dir, sign, g, tl, smem, stream);
      if (rc) return rc;
    } else {
      const int threads = 128;
      const int blocks = (g.volume + threads - 1) / threads;
      if (precision == MUGIQ_B200_PREC_DOUBLE)
        displace_site_kernel<double><<<blocks, threads, 0, stream>>>(batch, (const double *)gauge_d, dir, sign, g);
      else
        displace_site_kernel<float><<<blocks, threads, 0, stream>>>(batch, (const float *)gauge_d, dir, sign, g);
      MUGIQ_LAUNCH_CHECK();
    }
  }
  return MUGIQ_B200_OK;
}

int displace(void *dst_d, const void *src_d, const void *gauge_d, int dir, int sign, const LatGeom &g,
             int precision, cudaStream_t stream) {
  return displace_batch(&dst_d, &src_d, 1, gauge_d, dir, sign, g, precision, stream);
}

}  // namespace mugiq_b200
