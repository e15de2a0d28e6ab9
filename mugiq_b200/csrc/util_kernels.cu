// util_kernels.cu — stage 3 (gamma-basis / time-slice reorder), the phase matrix of stage 4, and the
// QUDA-order <-> site-major layout conversion.
//
// Replaces /root/reference/lib/mugiq_util_kernels.cu:3-35 (phaseMatrix_kernel), :59-99
// (convertIdxOrder_mapGamma_kernel) and their wrappers lib/contract_wrappers.cu:50-77,133-156.
#include "kernels.cuh"

namespace mugiq_b200 {

// ---------------------------------------------------------------------------------------------------
// Reorder:  out[t + Lt*((15-G) + 16*iL) + Lt*nData*v3] = sign[G] * in[x_eo + V4*(G + 16*iL)]
//
// Pure data movement (512 B per site and loop in FP64).  The input is contiguous in x_eo for fixed (G, iL): for a
// tile of R consecutive-y lattice rows the sites of one (t, parity) are ONE run of R*Lx/2 complex numbers; the output
// is contiguous in t for fixed (G', iL, v3).  A CTA owns one (G, iL) and one tile of rows for all t: it reads
// 2*Lt runs (a warp per run, lanes along the run), transposes through a padded shared tile indexed
// [parity][position in run][t] (conflict-free both ways: odd row stride, and the parity of a site alternates with t
// while its position in the run does not) and writes Lt consecutive complex numbers per site (a warp per site, lanes
// along t).  The reference writes each element with stride Lt*nData between neighbouring x
// (lib/mugiq_util_kernels.cu:92-96).
// ---------------------------------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256)
reorder_mapgamma_kernel(F *__restrict__ out, const F *__restrict__ in, const int nLoop, const int R, const LatGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cplx<F> *tile = reinterpret_cast<Cplx<F> *>(smem_raw);  // [2][R*Lh][Lt | 1]
  constexpr GammaTables gt = gamma_tables();
  const int Lt = g.L[3], Ltp = Lt | 1;
  const int NS = R * g.Lh;  // sites of the tile per parity
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int row0 = blockIdx.x * R;  // first row y + Ly*z of the tile
  const int idata = blockIdx.y, G = idata & 15, iL = idata >> 4;
  const int nData = 16 * nLoop;
  F sgn = 1;
  int Gp = 0;
#pragma unroll
  for (int q = 0; q < 16; q++)
    if (q == G) {
      sgn = (F)gt.map_sign[q];
      Gp = gt.map_index[q];
    }

  const F *src = in + 2 * ((size_t)g.volume * idata + (size_t)row0 * g.Lh);
  for (int tp = warp; tp < 2 * Lt; tp += nwarp) {
    const int t = tp >> 1, pty = tp & 1;
    const F *run = src + 2 * ((size_t)pty * g.volumeCB + (size_t)t * (g.V3 >> 1));
    for (int i = lane; i < NS; i += 32) {
      Cplx<F> z = ldg_c<F>(run + 2 * i);
      z.re *= sgn;
      z.im *= sgn;
      tile[(pty * NS + i) * Ltp + t] = z;
    }
  }
  __syncthreads();
  const int yz0 = (row0 % g.L[1]) + (row0 / g.L[1]);  // y + z of the first row (parity only)
  for (int ls = warp; ls < R * g.L[0]; ls += nwarp) {
    const int a = ls / g.L[0], x = ls - a * g.L[0];
    const int i = ls >> 1;                      // position in the run: (a*Lx + x) >> 1
    const int c = (x + yz0 + a) & 1;            // parity of the site at t = 0
    const size_t v3 = (size_t)(row0 + a) * g.L[0] + x;
    F *dst = out + 2 * ((size_t)Lt * (Gp + 16 * iL) + (size_t)Lt * nData * v3);
    for (int t = lane; t < Lt; t += 32) st_c<F>(dst + 2 * t, tile[((((c + t) & 1) * NS) + i) * Ltp + t]);
  }
}

int reorder_mapgamma(void *out_d, const void *in_d, int nLoop, const LatGeom &g, int precision, cudaStream_t stream) {
  // rows per tile: R | Ly with R*Lx/2 <= 32 (one warp-wide read per run), shrunk if the tile exceeds 64 KB
  const size_t per_row = (size_t)2 * g.Lh * (g.L[3] | 1) * 2 * prec_bytes(precision);
  int R = 1;
  for (int r = 1; r <= g.L[1]; r++)
    if (g.L[1] % r == 0 && r * g.Lh <= 32 && r * per_row <= 64 * 1024) R = r;
  const size_t smem = (size_t)R * per_row;
  if (smem > 200 * 1024) return set_error(MUGIQ_B200_EINVAL, "reorder_mapgamma: Lx*Lt = %d*%d too large for the tile", g.L[0], g.L[3]);
  const dim3 grid(g.L[1] * g.L[2] / R, 16 * nLoop);
  ProfScope prof(K_REORDER, stream, 2.0 * 16.0 * nLoop * (double)g.volume * 2.0 * prec_bytes(precision));
  if (precision == MUGIQ_B200_PREC_DOUBLE) {
    MUGIQ_CUDA_CHECK(cudaFuncSetAttribute(reorder_mapgamma_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    reorder_mapgamma_kernel<double><<<grid, 256, smem, stream>>>((double *)out_d, (const double *)in_d, nLoop, R, g);
  } else {
    MUGIQ_CUDA_CHECK(cudaFuncSetAttribute(reorder_mapgamma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    reorder_mapgamma_kernel<float><<<grid, 256, smem, stream>>>((float *)out_d, (const float *)in_d, nLoop, R, g);
  }
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Phase matrix: phase[v3 + V3*im] = cos(2 pi phi) + i*sgn*sin(2 pi phi)   (lib/mugiq_util_kernels.cu:23-31)
// phi is accumulated in double for both precisions and reduced with sincospi (exact argument
// reduction), which agrees with the reference's cos(2.0*PI*phase) to a few ulp.
// ---------------------------------------------------------------------------------------------------
struct PhaseArg {
  int localL[3], totalL[3], commCoord[3];
  int V3, Nmom, ftsign;
};

template <typename F>
__global__ void __launch_bounds__(256)
phase_matrix_kernel(F *__restrict__ phase, const int *__restrict__ mom, const PhaseArg a) {
  const int v3 = blockIdx.x * blockDim.x + threadIdx.x;
  const int im = blockIdx.y;
  if (v3 >= a.V3) return;
  const int a1 = v3 / a.localL[0];
  const int a2 = a1 / a.localL[1];
  const int gx = (v3 - a1 * a.localL[0]) + a.commCoord[0] * a.localL[0];
  const int gy = (a1 - a2 * a.localL[1]) + a.commCoord[1] * a.localL[1];
  const int gz = a2 + a.commCoord[2] * a.localL[2];
  const double phi = (double)(mom[0 + 3 * im] * gx) / (double)a.totalL[0] +
                     (double)(mom[1 + 3 * im] * gy) / (double)a.totalL[1] +
                     (double)(mom[2 + 3 * im] * gz) / (double)a.totalL[2];
  double s, c;
  sincospi(2.0 * phi, &s, &c);
  st_c<F>(phase + 2 * ((size_t)v3 + (size_t)a.V3 * im), make_c<F>((F)c, (F)((double)a.ftsign * s)));
}

// small persistent device buffer for the momentum table (avoids a cudaMalloc/cudaFree per call as in
// lib/contract_wrappers.cu:54-76)
int phase_matrix(void *phase_d, const int *mom_h, int Nmom, int ftsign, const int localL[4], const int totalL[4],
                 const int commCoord[4], int precision, cudaStream_t stream) {
  PhaseArg a;
  a.V3 = 1;
  for (int i = 0; i < 3; i++) {
    a.localL[i] = localL[i];
    a.totalL[i] = totalL[i];
    a.commCoord[i] = commCoord ? commCoord[i] : 0;
    a.V3 *= localL[i];
  }
  a.Nmom = Nmom;
  a.ftsign = ftsign;
  int *mom_d = nullptr;
  MUGIQ_CUDA_CHECK(cudaMallocAsync((void **)&mom_d, sizeof(int) * 3 * Nmom, stream));
  MUGIQ_CUDA_CHECK(cudaMemcpyAsync(mom_d, mom_h, sizeof(int) * 3 * Nmom, cudaMemcpyHostToDevice, stream));
  const dim3 grid((a.V3 + 255) / 256, Nmom);
  {
    ProfScope prof(K_PHASE, stream, (double)a.V3 * Nmom * 2.0 * prec_bytes(precision));
    if (precision == MUGIQ_B200_PREC_DOUBLE)
      phase_matrix_kernel<double><<<grid, 256, 0, stream>>>((double *)phase_d, mom_d, a);
    else
      phase_matrix_kernel<float><<<grid, 256, 0, stream>>>((float *)phase_d, mom_d, a);
    MUGIQ_LAUNCH_CHECK();
  }
  MUGIQ_CUDA_CHECK(cudaFreeAsync(mom_d, stream));
  // mom_h may be a temporary of the caller: make the copy complete before returning
  MUGIQ_CUDA_CHECK(cudaStreamSynchronize(stream));
  return MUGIQ_B200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Layout conversion between QUDA's native colour-spinor orders and the canonical site-major order.
// QUDA FloatNOrder<Float,4,3,N>: real element k = 2*(3*s+c)+reim of site x_cb lives at
// parity_offset + ((k/N)*stride + x_cb)*N + k%N with stride = volumeCB (no pad), parity_offset = parity*volumeCB*24,
// i.e. in complex units  FLOAT2: comp*Vh + x_cb        FLOAT4: (comp/2)*2*Vh + 2*x_cb + comp%2.
// A CTA converts kConvSites consecutive sites of one parity through a padded shared-memory tile, so that BOTH sides
// move in full 128-byte lines: the QUDA side as 12 (FLOAT2) or 6 (FLOAT4) contiguous runs, the site-major side as
// one contiguous block of kConvSites*192 B.  (The first version wrote 16 B per thread with a 192-B stride: 2.0 TB/s.)
// ---------------------------------------------------------------------------------------------------
constexpr int kConvSites = 64;
constexpr int kConvPad = kSpinorLen + 1;  // 13 complex per site: 208-B stride, conflict-free 16-B accesses

constexpr int kConvBatch = 64;  // fields per launch (blockIdx.y)
struct ConvBatch {
  void *dst[kConvBatch];
  const void *src[kConvBatch];
};

template <typename F>
__global__ void __launch_bounds__(256)
convert_spinor_kernel(const ConvBatch batch, const int order, const int to_site, const LatGeom g) {
  F *__restrict__ dst = static_cast<F *>(batch.dst[blockIdx.y]);
  const F *__restrict__ src = static_cast<const F *>(batch.src[blockIdx.y]);
  __shared__ __align__(16) unsigned char tile_raw[kConvSites * kConvPad * 2 * sizeof(F)];
  Cplx<F> *tile = reinterpret_cast<Cplx<F> *>(tile_raw);
  const int blocks_per_parity = (g.volumeCB + kConvSites - 1) / kConvSites;
  const int pty = blockIdx.x / blocks_per_parity;
  const int x0 = (blockIdx.x % blocks_per_parity) * kConvSites;
  const int nsite = min(kConvSites, g.volumeCB - x0);
  const size_t pbase = (size_t)pty * g.volumeCB * kSpinorLen;  // complex offset of this parity (both layouts)
  F *quda = to_site ? const_cast<F *>(src) : dst;
  F *site = to_site ? dst : const_cast<F *>(src);
  const int nelem = nsite * kSpinorLen;
  // element e of the QUDA-side traversal: contiguous runs of one plane
  auto quda_elem = [&](int e, int &s_loc, int &comp, size_t &off) {
    if (order == MUGIQ_B200_ORDER_FLOAT2) {
      comp = e / nsite;
      s_loc = e - comp * nsite;
      off = pbase + (size_t)comp * g.volumeCB + x0 + s_loc;
    } else {
      const int j = e / (2 * nsite), r = e - j * 2 * nsite;
      s_loc = r >> 1;
      comp = 2 * j + (r & 1);
      off = pbase + (size_t)j * 2 * g.volumeCB + 2 * (size_t)x0 + r;
    }
  };
  if (to_site) {
    for (int e = threadIdx.x; e < nelem; e += blockDim.x) {
      int s_loc, comp;
      size_t off;
      quda_elem(e, s_loc, comp, off);
      tile[s_loc * kConvPad + comp] = ldg_c<F>(quda + 2 * off);
    }
    __syncthreads();
    F *out = site + 2 * (pbase + (size_t)x0 * kSpinorLen);
    for (int e = threadIdx.x; e < nelem; e += blockDim.x) st_c<F>(out + 2 * e, tile[(e / kSpinorLen) * kConvPad + e % kSpinorLen]);
  } else {
    const F *in = site + 2 * (pbase + (size_t)x0 * kSpinorLen);
    for (int e = threadIdx.x; e < nelem; e += blockDim.x) tile[(e / kSpinorLen) * kConvPad + e % kSpinorLen] = ldg_c<F>(in + 2 * e);
    __syncthreads();
    for (int e = threadIdx.x; e < nelem; e += blockDim.x) {
      int s_loc, comp;
      size_t off;
      quda_elem(e, s_loc, comp, off);
      st_c<F>(quda + 2 * off, tile[s_loc * kConvPad + comp]);
    }
  }
}

int convert_spinor_batch(void *const *dst_d, const void *const *src_d, int n, int order, bool to_site, const LatGeom &g,
                         int precision, cudaStream_t stream) {
  const size_t total = (size_t)g.volume * kSpinorLen;
  const int blocks = 2 * ((g.volumeCB + kConvSites - 1) / kConvSites);
  for (int done = 0; done < n; done += kConvBatch) {
    ConvBatch batch;
    const int nb = n - done < kConvBatch ? n - done : kConvBatch;
    for (int i = 0; i < nb; i++) {
      batch.dst[i] = dst_d[done + i];
      batch.src[i] = src_d[done + i];
    }
    ProfScope prof(K_CONVERT, stream, 2.0 * (double)total * nb * 2.0 * prec_bytes(precision));
    const dim3 grid(blocks, nb);
    if (precision == MUGIQ_B200_PREC_DOUBLE)
      convert_spinor_kernel<double><<<grid, 256, 0, stream>>>(batch, order, to_site, g);
    else
      convert_spinor_kernel<float><<<grid, 256, 0, stream>>>(batch, order, to_site, g);
    MUGIQ_LAUNCH_CHECK();
  }
  return MUGIQ_B200_OK;
}

int convert_spinor(void *dst_d, const void *src_d, int order, bool to_site, const LatGeom &g, int precision,
                   cudaStream_t stream) {
  return convert_spinor_batch(&dst_d, &src_d, 1, order, to_site, g, precision, stream);
}

}  // namespace mugiq_b200
