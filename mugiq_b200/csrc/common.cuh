// common.cuh — geometry, complex arithmetic and gamma tables shared by all sm_100a kernels.
//
// Index conventions restate QUDA's public definitions used by the reference kernels
// (getCoords / linkIndexP1 / linkIndexM1 as called in /root/reference/lib/mugiq_displace_kernels.cu:128-168
// and lib/mugiq_util_kernels.cu:75); see include/mugiq_b200.h for the memory layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mugiq_b200.h"

namespace mugiq_b200 {

constexpr int kNs = 4;       // N_SPIN_   (include/util_mugiq.h:13)
constexpr int kNc = 3;       // N_COLOR_  (include/util_mugiq.h:14)
constexpr int kNg = 16;      // N_GAMMA_  (include/util_mugiq.h:15)
constexpr int kSpinorLen = 12;  // complex per site
constexpr int kLinkLen = 9;     // complex per link

// ---- host-side error plumbing (cabi.cu owns the storage) ------------------------------------------
int set_error(int code, const char *fmt, ...);
#define MUGIQ_CUDA_CHECK(expr)                                                                  \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::mugiq_b200::set_error(MUGIQ_B200_ECUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, \
                                     #expr, cudaGetErrorString(_e));                            \
  } while (0)
#define MUGIQ_LAUNCH_CHECK() MUGIQ_CUDA_CHECK(cudaGetLastError())

// ---- lattice geometry, passed to kernels by value --------------------------------------------------
struct LatGeom {
  int L[4];      // x,y,z,t
  int Lh;        // L[0]/2
  int volumeCB;  // sites per parity
  int volume;    // V4
  int V3;        // spatial volume
};

inline LatGeom make_geom(const int L[4]) {
  LatGeom g;
  for (int i = 0; i < 4; i++) g.L[i] = L[i];
  g.Lh = L[0] / 2;
  g.volume = L[0] * L[1] * L[2] * L[3];
  g.volumeCB = g.volume / 2;
  g.V3 = L[0] * L[1] * L[2];
  return g;
}

int check_geom(const mugiq_b200_geom_t *geom, const char *who);       // cabi.cu
int check_geom_even(const mugiq_b200_geom_t *geom, const char *who);  // cabi.cu: every extent even (displacements)

// QUDA getCoords(x, cb_index, X, parity) for a full (two-parity) field.
__host__ __device__ inline void get_coords(int x[4], int cb, int parity, const LatGeom &g) {
  const int za = cb / g.Lh;
  const int zb = za / g.L[1];
  x[1] = za - zb * g.L[1];
  x[3] = zb / g.L[2];
  x[2] = zb - x[3] * g.L[2];
  const int x1odd = (x[1] + x[2] + x[3] + parity) & 1;
  x[0] = 2 * cb + x1odd - za * g.L[0];
}

// checkerboard index of a site given by coordinates (QUDA linkIndex: lexicographic >> 1)
__host__ __device__ inline int cb_index(const int x[4], const LatGeom &g) {
  return (x[0] + g.L[0] * (x[1] + g.L[1] * (x[2] + g.L[2] * x[3]))) >> 1;
}

// ---- complex helpers ---------------------------------------------------------------------------------
template <typename F> struct Cplx {
  F re, im;
};
template <typename F> struct vec2_of;
template <> struct vec2_of<double> { using type = double2; };
template <> struct vec2_of<float> { using type = float2; };

template <typename F> __device__ __forceinline__ Cplx<F> make_c(F re, F im) {
  Cplx<F> c;
  c.re = re;
  c.im = im;
  return c;
}
// acc += a * b
template <typename F> __device__ __forceinline__ void cmac(Cplx<F> &acc, const Cplx<F> a, const Cplx<F> b) {
  acc.re = fma(a.re, b.re, acc.re);
  acc.re = fma(-a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im);
  acc.im = fma(a.im, b.re, acc.im);
}
// a * b
template <typename F> __device__ __forceinline__ Cplx<F> cmul(const Cplx<F> a, const Cplx<F> b) {
  Cplx<F> r;
  r.re = a.re * b.re;
  r.re = fma(-a.im, b.im, r.re);
  r.im = a.re * b.im;
  r.im = fma(a.im, b.re, r.im);
  return r;
}
// acc += conj(a) * b
template <typename F> __device__ __forceinline__ void cmac_conj(Cplx<F> &acc, const Cplx<F> a, const Cplx<F> b) {
  acc.re = fma(a.re, b.re, acc.re);
  acc.re = fma(a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im);
  acc.im = fma(-a.im, b.re, acc.im);
}

// 128-bit (FP64) / 64-bit (FP32) complex load through the read-only path.
template <typename F> __device__ __forceinline__ Cplx<F> ldg_c(const F *p) {
  using V = typename vec2_of<F>::type;
  const V v = __ldg(reinterpret_cast<const V *>(p));
  return make_c<F>(v.x, v.y);
}
template <typename F> __device__ __forceinline__ void st_c(F *p, const Cplx<F> c) {
  using V = typename vec2_of<F>::type;
  V v;
  v.x = c.re;
  v.y = c.im;
  *reinterpret_cast<V *>(p) = v;
}

// ---- gamma tables: DeGrand-Rossi basis, G(n) = g1^n0 g2^n1 g3^n2 g4^n3 ----------------------------------
// Values restate /root/reference/include/gamma.h:33-69 (one non-zero entry per row:
// G(n)_{ij} = RowValue[n][i] * (ColumnIndex[n][i] == j)).  Each value is one of +1, -1, +i, -i and is
// stored as a power of i: 0 -> +1, 1 -> +i, 2 -> -1, 3 -> -i, so that kernels apply it with adds and
// swaps only ("FP64 ALU budget").
struct GammaTables {
  int8_t ipow[16][4];
  int8_t col[16][4];
  int8_t map_sign[16];   // +1 / -1   (include/gamma.h:99-102: minus for {3,6,9,11,12,14})
  int8_t map_index[16];  // 15 - G    (include/gamma.h:105-109)
};

__host__ __device__ constexpr GammaTables gamma_tables() {
  return GammaTables{
      {{0, 0, 0, 0},   // G0  = 1
       {1, 1, 3, 3},   // G1  = g1
       {2, 0, 0, 2},   // G2  = g2
       {3, 1, 3, 1},   // G3  = g1g2
       {1, 3, 3, 1},   // G4  = g3
       {2, 0, 2, 0},   // G5  = g1g3
       {3, 3, 3, 3},   // G6  = g2g3
       {0, 0, 2, 2},   // G7  = g1g2g3   =  g5g4
       {0, 0, 0, 0},   // G8  = g4
       {1, 1, 3, 3},   // G9  = g1g4
       {2, 0, 0, 2},   // G10 = g2g4
       {3, 1, 3, 1},   // G11 = g1g2g4   = -g5g3
       {1, 3, 3, 1},   // G12 = g3g4
       {2, 0, 2, 0},   // G13 = g1g3g4   =  g5g2
       {3, 3, 3, 3},   // G14 = g2g3g4   = -g5g1
       {0, 0, 2, 2}},  // G15 = g1g2g3g4 =  g5
      {{0, 1, 2, 3},
       {3, 2, 1, 0},
       {3, 2, 1, 0},
       {0, 1, 2, 3},
       {2, 3, 0, 1},
       {1, 0, 3, 2},
       {1, 0, 3, 2},
       {2, 3, 0, 1},
       {2, 3, 0, 1},
       {1, 0, 3, 2},
       {1, 0, 3, 2},
       {2, 3, 0, 1},
       {0, 1, 2, 3},
       {3, 2, 1, 0},
       {3, 2, 1, 0},
       {0, 1, 2, 3}},
      {1, 1, 1, -1, 1, 1, -1, 1, 1, -1, 1, -1, -1, 1, -1, 1},
      {15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0}};
}

// acc += i^p * z, p a compile-time constant after unrolling
template <typename F> __device__ __forceinline__ void add_ipow(Cplx<F> &acc, const int p, const Cplx<F> z) {
  if (p == 0) {
    acc.re += z.re;
    acc.im += z.im;
  } else if (p == 1) {
    acc.re -= z.im;
    acc.im += z.re;
  } else if (p == 2) {
    acc.re -= z.re;
    acc.im -= z.im;
  } else {
    acc.re += z.im;
    acc.im -= z.re;
  }
}

// Gamma projection of a 4x4 spin matrix M[be][al] = sum_c conj(vL[be,c]) vR[al,c]:
// T[G] = sum_{s2} rowval[G][s2] * M[s2][col[G][s2]]   (lib/mugiq_contract_kernels.cu:111-117)
template <typename F> __device__ __forceinline__ void gamma_project(Cplx<F> T[16], const Cplx<F> M[4][4]) {
  constexpr GammaTables gt = gamma_tables();
#pragma unroll
  for (int G = 0; G < 16; G++) {
    Cplx<F> t = make_c<F>(0, 0);
#pragma unroll
    for (int s2 = 0; s2 < 4; s2++) add_ipow(t, gt.ipow[G][s2], M[s2][gt.col[G][s2]]);
    T[G] = t;
  }
}

}  // namespace mugiq_b200
