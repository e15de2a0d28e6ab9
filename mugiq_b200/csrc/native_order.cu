// native_order.cu — stage 1 and stage 2 directly on fields in QUDA's native FLOAT2 / FLOAT4 orders.
//
// The reference's kernels read eigenvectors through QUDA's FieldOrderCB accessors
// (/root/reference/lib/mugiq_contract_kernels.cu:82-83, lib/mugiq_displace_kernels.cu:85-113), i.e. in the order the
// eigensolver left them in.  The canonical order of this library is site-major (what the fused kernel's bulk-TMA tiles
// want), and a caller holding FLOAT2 / FLOAT4 fields reaches it through mugiq_b200_ingest_spinor.  For the
// reference-shaped SINGLE calls (performLoopContraction, performCovariantDisplacementVector) that conversion would cost
// more than the call itself, so these two kernels work on the native orders in place: in FLOAT2
// ([parity][spin*3+colour][x_cb]) and FLOAT4 ([parity][j][x_cb][2]) consecutive sites of one component are adjacent, so
// thread = site gives fully coalesced 128-bit accesses without any staging, the source AND the destination of a
// displacement may be native-order fields, and no scratch field is allocated.
#include <algorithm>

#include "kernels.cuh"

namespace mugiq_b200 {

// complex index of component j = 3*spin + colour at (parity, x_cb)   (oracle/quda_shim/quda_shim_core.h restates the
// same two formulas for the reference's accessors; tests/test_ref_kernels.py runs both on the same buffers)
template <int ORDER> __device__ __forceinline__ size_t native_index(int parity, int x_cb, int j, int volumeCB) {
  const size_t off = (size_t)parity * kSpinorLen * volumeCB;
  if (ORDER == MUGIQ_B200_ORDER_FLOAT2) return off + (size_t)j * volumeCB + x_cb;
  return off + ((size_t)(j >> 1) * volumeCB + x_cb) * 2 + (j & 1);
}

template <typename F, int ORDER, bool kSame>
__global__ void __launch_bounds__(128)
contract_native_kernel(F *__restrict__ loop, const VecBatch batch, const int accumulate, const LatGeom g) {
  const int x_eo = blockIdx.x * blockDim.x + threadIdx.x;
  if (x_eo >= g.volume) return;
  const int pty = x_eo >= g.volumeCB, x_cb = x_eo - pty * g.volumeCB;
  Cplx<F> M[4][4];
#pragma unroll
  for (int be = 0; be < 4; be++)
#pragma unroll
    for (int al = 0; al < 4; al++) M[be][al] = make_c<F>(0, 0);
  for (int n = 0; n < batch.nvec; n++) {
    const F *pl = static_cast<const F *>(batch.vL[n]), *pr = static_cast<const F *>(batch.vR[n]);
    const F inv_sigma = (F)batch.inv_sigma[n];
    Cplx<F> l[kSpinorLen], r[kSpinorLen];
#pragma unroll
    for (int j = 0; j < kSpinorLen; j++) {
      const size_t i = native_index<ORDER>(pty, x_cb, j, g.volumeCB);
      l[j] = ldg_c<F>(pl + 2 * i);
      r[j] = kSame ? l[j] : ldg_c<F>(pr + 2 * i);
    }
#pragma unroll
    for (int j = 0; j < kSpinorLen; j++) {
      l[j].re *= inv_sigma;
      l[j].im *= inv_sigma;
    }
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

struct NativeDispBatch {
  const void *src[16];
  void *dst[16];
  int nvec;
};

template <typename F, int ORDER>
__global__ void __launch_bounds__(128)
displace_native_kernel(const NativeDispBatch batch, const F *__restrict__ gauge, const int dir, const int sign, const LatGeom g) {
  const int x_eo = blockIdx.x * blockDim.x + threadIdx.x;
  if (x_eo >= g.volume) return;
  const int pty = x_eo >= g.volumeCB, x_cb = x_eo - pty * g.volumeCB;
  int x[4];
  get_coords(x, x_cb, pty, g);
  x[dir] = sign ? (x[dir] + 1 == g.L[dir] ? 0 : x[dir] + 1) : (x[dir] == 0 ? g.L[dir] - 1 : x[dir] - 1);
  const int nbr_cb = cb_index(x, g);
  const size_t link_site = sign ? (size_t)x_eo : (size_t)nbr_cb + (size_t)(1 - pty) * g.volumeCB;
  const F *pu = gauge + 2 * kLinkLen * ((size_t)dir * g.volume + link_site);
  Cplx<F> U[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const Cplx<F> u = ldg_c<F>(pu + 2 * (r * 3 + c));
      if (sign)
        U[r][c] = u;
      else
        U[c][r] = make_c<F>(u.re, -u.im);
    }
  for (int n = 0; n < batch.nvec; n++) {
    const F *ps = static_cast<const F *>(batch.src[n]);
    F *pd = static_cast<F *>(batch.dst[n]);
    Cplx<F> v[kSpinorLen];
#pragma unroll
    for (int j = 0; j < kSpinorLen; j++) v[j] = ldg_c<F>(ps + 2 * native_index<ORDER>(1 - pty, nbr_cb, j, g.volumeCB));
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        Cplx<F> acc = cmul(U[c][0], v[s * 3 + 0]);
        cmac(acc, U[c][1], v[s * 3 + 1]);
        cmac(acc, U[c][2], v[s * 3 + 2]);
        st_c<F>(pd + 2 * native_index<ORDER>(pty, x_cb, s * 3 + c, g.volumeCB), acc);
      }
  }
}

template <typename F, int ORDER>
static int launch_contract_native(void *loop_d, const VecBatch &b, bool same, int accumulate, const LatGeom &g, cudaStream_t stream) {
  const int blocks = (g.volume + 127) / 128;
  const double S = kSpinorLen * 2.0 * sizeof(F), A = 16 * 2.0 * sizeof(F);
  ProfScope prof(K_CONTRACT, stream, (double)g.volume * (b.nvec * (same ? S : 2 * S) + (accumulate ? 2 * A : A)),
                 (double)g.volume * b.nvec * 2.0 * (192 + 24));
  if (same)
    contract_native_kernel<F, ORDER, true><<<blocks, 128, 0, stream>>>((F *)loop_d, b, accumulate, g);
  else
    contract_native_kernel<F, ORDER, false><<<blocks, 128, 0, stream>>>((F *)loop_d, b, accumulate, g);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

int contract_batch_native(void *loop_d, const void *const *vL, const void *const *vR, const double *sigma, int nvec, int order,
                          int accumulate, const LatGeom &g, int precision, cudaStream_t stream) {
  for (int done = 0; done < nvec;) {
    VecBatch b;
    b.nvec = std::min(nvec - done, kMaxBatch);
    for (int i = 0; i < b.nvec; i++) {
      b.vL[i] = vL[done + i];
      b.vR[i] = vR ? vR[done + i] : vL[done + i];
      b.inv_sigma[i] = inv_sigma_of(sigma[done + i], precision);
    }
    const int acc = accumulate || done > 0;
    const bool same = vR == nullptr;
    int rc;
    if (precision == MUGIQ_B200_PREC_DOUBLE)
      rc = order == MUGIQ_B200_ORDER_FLOAT2 ? launch_contract_native<double, MUGIQ_B200_ORDER_FLOAT2>(loop_d, b, same, acc, g, stream)
                                            : launch_contract_native<double, MUGIQ_B200_ORDER_FLOAT4>(loop_d, b, same, acc, g, stream);
    else
      rc = order == MUGIQ_B200_ORDER_FLOAT2 ? launch_contract_native<float, MUGIQ_B200_ORDER_FLOAT2>(loop_d, b, same, acc, g, stream)
                                            : launch_contract_native<float, MUGIQ_B200_ORDER_FLOAT4>(loop_d, b, same, acc, g, stream);
    if (rc) return rc;
    done += b.nvec;
  }
  return MUGIQ_B200_OK;
}

int displace_batch_native(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir, int sign,
                          int order, const LatGeom &g, int precision, cudaStream_t stream) {
  const int blocks = (g.volume + 127) / 128;
  for (int done = 0; done < nvec; done += 16) {
    NativeDispBatch b;
    b.nvec = std::min(nvec - done, 16);
    for (int i = 0; i < b.nvec; i++) {
      b.src[i] = src_d[done + i];
      b.dst[i] = dst_d[done + i];
    }
    const double pb = (double)prec_bytes(precision);
    ProfScope prof(K_DISPLACE, stream, (double)g.volume * pb * 2.0 * (b.nvec * 2.0 * kSpinorLen + kLinkLen),
                   (double)g.volume * b.nvec * 288.0);
#define MUGIQ_DISP_NATIVE(F, ORD) \
  displace_native_kernel<F, ORD><<<blocks, 128, 0, stream>>>(b, (const F *)gauge_d, dir, sign, g)
    if (precision == MUGIQ_B200_PREC_DOUBLE) {
      if (order == MUGIQ_B200_ORDER_FLOAT2)
        MUGIQ_DISP_NATIVE(double, MUGIQ_B200_ORDER_FLOAT2);
      else
        MUGIQ_DISP_NATIVE(double, MUGIQ_B200_ORDER_FLOAT4);
    } else {
      if (order == MUGIQ_B200_ORDER_FLOAT2)
        MUGIQ_DISP_NATIVE(float, MUGIQ_B200_ORDER_FLOAT2);
      else
        MUGIQ_DISP_NATIVE(float, MUGIQ_B200_ORDER_FLOAT4);
    }
#undef MUGIQ_DISP_NATIVE
    MUGIQ_LAUNCH_CHECK();
  }
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200
