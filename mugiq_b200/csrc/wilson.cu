// wilson.cu — gauge-only and loop-buffer-only helper kernels of the fused loop schedule (loop_fused.cu).
//
// The reference applies a displacement of length k as k successive one-link hops of every eigenvector
// (/root/reference/lib/loop_mugiq.cpp:489-491 -> lib/displace.cpp:55-67 -> lib/mugiq_displace_kernels.cu:156-185):
//   plus : R_k(x) = U(x) U(x+mu) ... U(x+(k-1)mu) v(x+k mu)            =: W+_k(x) v(x+k mu)
//   minus: R_k(x) = U(x-mu)^dag U(x-2mu)^dag ... U(x-k mu)^dag v(x-k mu) =: W-_k(x) v(x-k mu),  W-_k(x) = [W+_k(x-k mu)]^dag
// The Wilson lines W depend on the gauge field only, so they are built once (here) and the eigenvector loop
// applies a single 3x3 multiplication per displaced loop.
//
// loop_minus_from_plus uses the exact identity (any links, no unitarity needed)
//   M-_k(x)[be][al] = conj( M+_k(x - k mu)[al][be] )   =>   T-_G(x) = h_G conj( T+_G(x - k mu) ),  Gamma_G^dag = h_G Gamma_G
// so that the minus-direction loops of an entry pair (+mu, -mu) cost one pass over the loop buffer instead of
// a second eigenvector sweep.
#include "fused.cuh"

namespace mugiq_b200 {

// full-site index of x + shift*dir (periodic) for the site x_eo
__device__ __forceinline__ int shifted_site(int x_eo, int dir, int shift, const LatGeom &g) {
  const int pty = x_eo >= g.volumeCB;
  const int x_cb = x_eo - pty * g.volumeCB;
  int x[4];
  get_coords(x, x_cb, pty, g);
  int v = (x[dir] + shift) % g.L[dir];
  if (v < 0) v += g.L[dir];
  x[dir] = v;
  const int npty = (pty + (shift & 1)) & 1;
  return cb_index(x, g) + npty * g.volumeCB;
}

template <typename F> __device__ __forceinline__ void load_link(Cplx<F> U[3][3], const F *p) {
#pragma unroll
  for (int k = 0; k < 9; k++) U[k / 3][k % 3] = ldg_c<F>(p + 2 * k);
}

template <typename F>
__global__ void __launch_bounds__(128)
wilson_extend_kernel(F *__restrict__ Wout, const F *__restrict__ Win, const F *__restrict__ gauge, const int dir,
                     const int shift, const LatGeom g) {
  const int x_eo = blockIdx.x * blockDim.x + threadIdx.x;
  if (x_eo >= g.volume) return;
  const int y_eo = shifted_site(x_eo, dir, shift, g);
  Cplx<F> A[3][3], B[3][3];
  load_link(A, Win + (size_t)x_eo * 18);
  load_link(B, gauge + ((size_t)dir * g.volume + y_eo) * 18);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      Cplx<F> acc = make_c<F>(0, 0);
#pragma unroll
      for (int k = 0; k < 3; k++) cmac(acc, A[r][k], B[k][c]);
      st_c<F>(Wout + (size_t)x_eo * 18 + 2 * (r * 3 + c), acc);
    }
}

template <typename F>
__global__ void __launch_bounds__(128)
wilson_minus_kernel(F *__restrict__ Wm, const F *__restrict__ Wp, const int dir, const int len, const LatGeom g) {
  const int x_eo = blockIdx.x * blockDim.x + threadIdx.x;
  if (x_eo >= g.volume) return;
  const int y_eo = shifted_site(x_eo, dir, -len, g);
  Cplx<F> A[3][3];
  load_link(A, Wp + (size_t)y_eo * 18);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) st_c<F>(Wm + (size_t)x_eo * 18 + 2 * (r * 3 + c), make_c<F>(A[c][r].re, -A[c][r].im));
}

// h_G with Gamma_G^dagger = h_G Gamma_G, from the compiled-in tables: (Gamma^dag)[a][b] = conj(Gamma[b][a])
__host__ __device__ constexpr int gamma_herm_sign(int G) {
  const GammaTables gt = gamma_tables();
  // row 0 of Gamma: value i^p0 at column c0.  Row 0 of Gamma^dag: conj of Gamma[b][0], b the row with col[b] == 0.
  int b = 0;
  for (int s = 0; s < 4; s++)
    if (gt.col[G][s] == 0) b = s;
  // Gamma^dag[0][b] = conj(i^ipow[b]) = i^(-ipow[b]);  h * Gamma[0][c0] with c0 == b for these matrices
  const int e = ((-gt.ipow[G][b] - gt.ipow[G][0]) % 4 + 4) % 4;  // h = i^e, e in {0, 2}
  return e == 0 ? 1 : -1;
}

// One launch derives every minus loop of a plan (blockIdx.y = derivation): the loop buffer is streamed once at full width
// instead of one 16-us launch per loop.
template <typename F>
__global__ void __launch_bounds__(256)
loop_minus_from_plus_kernel(F *__restrict__ pos, const MinusBatch b, const int accumulate, const LatGeom g) {
  const int x_eo = blockIdx.x * blockDim.x + threadIdx.x;
  if (x_eo >= g.volume) return;
  const MinusBatch::Item it = b.item[blockIdx.y];
  const int y_eo = shifted_site(x_eo, it.dir, -it.len, g);
  const F *plus = pos + 2 * (size_t)it.src * 16 * (size_t)g.volume;
  F *minus = pos + 2 * (size_t)it.dst * 16 * (size_t)g.volume;
#pragma unroll
  for (int G = 0; G < 16; G++) {
    const F h = (F)gamma_herm_sign(G);
    const Cplx<F> z = ldg_c<F>(plus + 2 * ((size_t)y_eo + (size_t)g.volume * G));
    F *po = minus + 2 * ((size_t)x_eo + (size_t)g.volume * G);
    Cplx<F> o = make_c<F>(h * z.re, -h * z.im);
    if (accumulate) {
      const Cplx<F> old = ldg_c<F>(po);
      o.re += old.re;
      o.im += old.im;
    }
    st_c<F>(po, o);
  }
}

int wilson_extend(void *Wout_d, const void *Win_d, const void *gauge_d, int dir, int shift, const LatGeom &g,
                  int precision, cudaStream_t stream) {
  const int blocks = (g.volume + 127) / 128;
  ProfScope prof(K_WILSON_LINE, stream, (double)g.volume * 3 * 18.0 * prec_bytes(precision));
  if (precision == MUGIQ_B200_PREC_DOUBLE)
    wilson_extend_kernel<double><<<blocks, 128, 0, stream>>>((double *)Wout_d, (const double *)Win_d,
                                                             (const double *)gauge_d, dir, shift, g);
  else
    wilson_extend_kernel<float><<<blocks, 128, 0, stream>>>((float *)Wout_d, (const float *)Win_d, (const float *)gauge_d,
                                                            dir, shift, g);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

int wilson_minus_from_plus(void *Wminus_d, const void *Wplus_d, int dir, int len, const LatGeom &g, int precision,
                           cudaStream_t stream) {
  const int blocks = (g.volume + 127) / 128;
  ProfScope prof(K_WILSON_LINE, stream, (double)g.volume * 2 * 18.0 * prec_bytes(precision));
  if (precision == MUGIQ_B200_PREC_DOUBLE)
    wilson_minus_kernel<double><<<blocks, 128, 0, stream>>>((double *)Wminus_d, (const double *)Wplus_d, dir, len, g);
  else
    wilson_minus_kernel<float><<<blocks, 128, 0, stream>>>((float *)Wminus_d, (const float *)Wplus_d, dir, len, g);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

int loop_minus_from_plus(void *dataPos_d, const MinusBatch &batch, int accumulate, const LatGeom &g, int precision,
                         cudaStream_t stream) {
  if (batch.n < 1) return MUGIQ_B200_OK;
  const dim3 blocks((g.volume + 255) / 256, batch.n);
  ProfScope prof(K_MINUS_FROM_PLUS, stream, (double)batch.n * g.volume * (accumulate ? 3 : 2) * 32.0 * prec_bytes(precision));
  if (precision == MUGIQ_B200_PREC_DOUBLE)
    loop_minus_from_plus_kernel<double><<<blocks, 256, 0, stream>>>((double *)dataPos_d, batch, accumulate, g);
  else
    loop_minus_from_plus_kernel<float><<<blocks, 256, 0, stream>>>((float *)dataPos_d, batch, accumulate, g);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200
