// mugiq_oracle.cpp — CPU restatement of MuGiq's disconnected-loop hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under mugiq_b200/ may import, link or execute this file; it is
// used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the
// checker and as the timed host baseline, never as the product path.
//
// *** PARITY: PINNED TO THE REFERENCE'S KERNEL CODE, QUDA'S ACCESSORS ASSUMED. ***  The reference (ckallidonis/mugiq)
// ships no unit tests, golden vectors or fixtures for this path, and its full path cannot be
// built here: QUDA is neither vendored nor installed (version unpinned: CMake takes MUGIQ_QUDA_HOME only,
// CMakeLists.txt:112-114).  What CAN be built is the reference's own kernel layer: lib/contract_wrappers.cu and
// lib/mugiq_{contract,displace,util}_kernels.cu compile UNMODIFIED against a small restatement of the QUDA interfaces
// they touch (oracle/quda_shim/, `make -C oracle ref` -> oracle/_ref/libmugiq_ref.so).  This file is checked against
//   (a) golden vectors those reference kernels produced on a B200 (tests/golden/ref_kernels_4x4x4x8.npz, generator
//       tests/golden/make_ref_golden.py; CPU suite tests/test_golden_ref.py),
//   (b) the same kernels live on the GPU box (tests/test_ref_kernels.py, which also compares the product's kernels
//       with them directly),
//   (c) the algebraic known-answer tests of tests/test_oracle_kats.py and an independent numpy
//       restatement (oracle/numpy_check.py).
// So the kernel arithmetic, the gamma tables and their assembly, which neighbour / link / parity / dagger a displacement
// picks, the reorder index and sign map and the phase formula are pinned to the reference's code; the QUDA semantics
// A1-A5 below (shared by this file and by the shim) remain assumptions, restated from QUDA's public definitions.
//
// QUDA semantics assumed (github.com/lattice/quda, develop of early 2020):
//  A1 getCoords(x, cb, X, parity): za=cb/(X0/2); zb=za/X1; x1=za-zb*X1; x3=zb/X2; x2=zb-x3*X2;
//     x0=2*cb+((x1+x2+x3+parity)&1)-za*X0                               (index_helper.cuh)
//  A2 linkIndex / linkIndexP1 / linkIndexM1 / linkIndexShift: lexicographic index of the periodically
//     wrapped coordinate, >>1; the neighbour has the opposite parity.
//  A3 colour-spinor accessor F(parity, x_cb, s, c); canonical storage here is site-major
//     [parity][x_cb][s][c] complex (component c + 3*s, include/util_mugiq.h:19).
//  A4 gauge accessor U(dir, x_cb, parity) returns the 3x3 matrix (row, col) of host QDP order
//     [dir][parity][x_cb][row][col]; conj(Matrix) is the Hermitian conjugate; Link*Vector is
//     y(s,c) = sum_c' U(c,c') x(s,c').
//  A5 a single process (comm_dim = 1, comm_coord = 0): the extended gauge field has border 0
//     (lib/displace.cpp:16) and no ghost exchange is needed.
//
// Each function cites the reference lines it follows.  Build: oracle/Makefile (g++ -O3 -fopenmp).
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int N_SPIN = 4, N_COLOR = 3, N_GAMMA = 16;

// include/gamma.h:32-49  (row value, as {re, im})
const int kRowValue[N_GAMMA][N_SPIN][2] = {
    {{1, 0}, {1, 0}, {1, 0}, {1, 0}},       {{0, 1}, {0, 1}, {0, -1}, {0, -1}},  {{-1, 0}, {1, 0}, {1, 0}, {-1, 0}},
    {{0, -1}, {0, 1}, {0, -1}, {0, 1}},     {{0, 1}, {0, -1}, {0, -1}, {0, 1}},  {{-1, 0}, {1, 0}, {-1, 0}, {1, 0}},
    {{0, -1}, {0, -1}, {0, -1}, {0, -1}},   {{1, 0}, {1, 0}, {-1, 0}, {-1, 0}},  {{1, 0}, {1, 0}, {1, 0}, {1, 0}},
    {{0, 1}, {0, 1}, {0, -1}, {0, -1}},     {{-1, 0}, {1, 0}, {1, 0}, {-1, 0}},  {{0, -1}, {0, 1}, {0, -1}, {0, 1}},
    {{0, 1}, {0, -1}, {0, -1}, {0, 1}},     {{-1, 0}, {1, 0}, {-1, 0}, {1, 0}},  {{0, -1}, {0, -1}, {0, -1}, {0, -1}},
    {{1, 0}, {1, 0}, {-1, 0}, {-1, 0}}};
// include/gamma.h:53-70
const int kColumnIdx[N_GAMMA][N_SPIN] = {{0, 1, 2, 3}, {3, 2, 1, 0}, {3, 2, 1, 0}, {0, 1, 2, 3}, {2, 3, 0, 1}, {1, 0, 3, 2},
                                         {1, 0, 3, 2}, {2, 3, 0, 1}, {2, 3, 0, 1}, {1, 0, 3, 2}, {1, 0, 3, 2}, {2, 3, 0, 1},
                                         {0, 1, 2, 3}, {3, 2, 1, 0}, {3, 2, 1, 0}, {0, 1, 2, 3}};
// include/gamma.h:99-102
const int kMinusGamma[6] = {3, 6, 9, 11, 12, 14};

struct Geom {
  int L[4];
  int volume, volumeCB, V3;
  explicit Geom(const int *l) {
    for (int i = 0; i < 4; i++) L[i] = l[i];
    V3 = L[0] * L[1] * L[2];
    volume = V3 * L[3];
    volumeCB = volume / 2;
  }
};

// A1
inline void get_coords(int x[4], int cb, const int X[4], int parity) {
  const int za = cb / (X[0] >> 1);
  const int zb = za / X[1];
  x[1] = za - zb * X[1];
  x[3] = zb / X[2];
  x[2] = zb - x[3] * X[2];
  const int x1odd = (x[1] + x[2] + x[3] + parity) & 1;
  x[0] = 2 * cb + x1odd - za * X[0];
}
// A2
inline int link_index(const int x[4], const int X[4]) { return (x[0] + X[0] * (x[1] + X[1] * (x[2] + X[2] * x[3]))) >> 1; }
inline int link_index_shift(const int x[4], int dir, int shift, const int X[4]) {
  int y[4] = {x[0], x[1], x[2], x[3]};
  y[dir] = ((y[dir] + shift) % X[dir] + X[dir]) % X[dir];
  return link_index(y, X);
}

template <typename F> using C = std::complex<F>;

// lib/contract_wrappers.cu:26-43: sign[] from minusGamma(), index[i] = 15 - i
void gamma_map(double sign[16], int index[16]) {
  for (int i = 0; i < 16; i++) {
    sign[i] = 1.0;
    index[i] = N_GAMMA - i - 1;
  }
  for (int g : kMinusGamma) sign[g] = -1.0;
}

// loopContract_kernel, lib/mugiq_contract_kernels.cu:98-120, for every site:
//   resG[be][al] = sum_c conj(vL[be,c]) vR[al,c];  trace_G = sum_s2 rowval[G][s2] resG[s2][col[G][s2]];
//   loop[x_eo + V*G] += inv_sigma * trace_G,  inv_sigma = 1.0/sigma (include/contract_util.cuh:133)
template <typename F>
void contract(C<F> *loop, const C<F> *vL, const C<F> *vR, F sigma, const Geom &g) {
  const F inv_sigma = (F)(1.0 / sigma);
#pragma omp parallel for schedule(static)
  for (int x = 0; x < g.volume; x++) {
    const C<F> *l = vL + (size_t)x * 12, *r = vR + (size_t)x * 12;
    C<F> resG[4][4];
    for (int be = 0; be < 4; be++)
      for (int al = 0; al < 4; al++) {
        C<F> s = 0;
        for (int kc = 0; kc < N_COLOR; kc++) s += std::conj(l[kc + 3 * be]) * r[kc + 3 * al];
        resG[be][al] = s;
      }
    for (int iG = 0; iG < N_GAMMA; iG++) {
      C<F> trace = 0;
      for (int s2 = 0; s2 < N_SPIN; s2++) {
        const int s1 = kColumnIdx[iG][s2];
        trace += C<F>((F)kRowValue[iG][s2][0], (F)kRowValue[iG][s2][1]) * resG[s2][s1];
      }
      loop[(size_t)x + (size_t)g.volume * iG] += inv_sigma * trace;
    }
  }
}

// covariantDisplacementVector_kernel, lib/mugiq_displace_kernels.cu:156-185 with getNbrSiteVec (:116-151)
// and getNbrLinkExtG (:39-74):  plus: dst(x) = U_d(x) src(x+d);  minus: dst(x) = U_d(x-d)^dag src(x-d)
template <typename F>
void displace(C<F> *dst, const C<F> *src, const C<F> *gauge, int dir, int sign, const Geom &g) {
#pragma omp parallel for schedule(static)
  for (int xeo = 0; xeo < g.volume; xeo++) {
    const int pty = xeo / g.volumeCB, x_cb = xeo % g.volumeCB;
    int coord[4];
    get_coords(coord, x_cb, g.L, pty);
    const int nbrPty = 1 - pty;
    const int nbrIdx = link_index_shift(coord, dir, sign ? +1 : -1, g.L);
    const C<F> *v = src + ((size_t)nbrPty * g.volumeCB + nbrIdx) * 12;
    C<F> U[3][3];
    if (sign) {
      const C<F> *u = gauge + (((size_t)dir * 2 + pty) * g.volumeCB + x_cb) * 9;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) U[r][c] = u[r * 3 + c];
    } else {
      const C<F> *u = gauge + (((size_t)dir * 2 + nbrPty) * g.volumeCB + nbrIdx) * 9;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) U[r][c] = std::conj(u[c * 3 + r]);
    }
    C<F> *d = dst + (size_t)xeo * 12;
    for (int s = 0; s < N_SPIN; s++)
      for (int c = 0; c < N_COLOR; c++) {
        C<F> acc = 0;
        for (int cp = 0; cp < N_COLOR; cp++) acc += U[c][cp] * v[cp + 3 * s];
        d[c + 3 * s] = acc;
      }
  }
}

// convertIdxOrder_mapGamma_kernel, lib/mugiq_util_kernels.cu:59-99
template <typename F>
void reorder_mapgamma(C<F> *out, const C<F> *in, int nLoop, const Geom &g) {
  double sign[16];
  int index[16];
  gamma_map(sign, index);
  const int nData = N_GAMMA * nLoop;
  const int Lx = g.L[0], Ly = g.L[1], Lt = g.L[3];
#pragma omp parallel for schedule(static)
  for (int tid = 0; tid < g.volume; tid++) {
    const int pty = tid / g.volumeCB, x_cb = tid % g.volumeCB;
    int crd[4];
    get_coords(crd, x_cb, g.L, pty);
    const int v3 = crd[0] + Lx * crd[1] + Lx * Ly * crd[2];
    const int t = crd[3];
    for (int ig = 0; ig < N_GAMMA; ig++)
      for (int iL = 0; iL < nLoop; iL++) {
        const size_t idxFrom = (size_t)tid + (size_t)g.volume * (ig + N_GAMMA * iL);
        const size_t idataTo = index[ig] + N_GAMMA * iL;
        const size_t idxTo = t + (size_t)Lt * idataTo + (size_t)Lt * nData * v3;
        out[idxTo] = (F)sign[ig] * in[idxFrom];
      }
  }
}

// phaseMatrix_kernel, lib/mugiq_util_kernels.cu:3-35 (PI = 2.0*asin(1.0), include/util_mugiq.h:7)
template <typename F>
void phase_matrix(C<F> *phase, const int *mom, int Nmom, int ftsign, const int *localL, const int *totalL,
                  const int *commCoord) {
  const double PI = 2.0 * asin(1.0);
  const int V3 = localL[0] * localL[1] * localL[2];
  for (int tid = 0; tid < V3; tid++) {
    const int a1 = tid / localL[0], a2 = a1 / localL[1];
    int gc[3] = {tid - a1 * localL[0] + commCoord[0] * localL[0], a1 - a2 * localL[1] + commCoord[1] * localL[1],
                 a2 + commCoord[2] * localL[2]};
    for (int im = 0; im < Nmom; im++) {
      F ph = 0;
      for (int id = 0; id < 3; id++) ph += mom[id + 3 * im] * gc[id] / (F)totalL[id];
      phase[(size_t)tid + (size_t)V3 * im] = C<F>((F)cos(2.0 * PI * ph), (F)((F)ftsign * sin(2.0 * PI * ph)));
    }
  }
}

// cublasZgemm(N, N, M, N, K, 1, A, M, B, K, 0, C, M), lib/loop_mugiq.cpp:364-377: naive column-major GEMM,
// accumulated in long double to serve as the accuracy reference
template <typename F>
void gemm(C<F> *Cm, const C<F> *A, const C<F> *B, long long M, int N, long long K) {
#pragma omp parallel for schedule(static)
  for (long long m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      long double re = 0, im = 0;
      for (long long k = 0; k < K; k++) {
        const C<F> a = A[m + M * k], b = B[k + K * n];
        re += (long double)a.real() * b.real() - (long double)a.imag() * b.imag();
        im += (long double)a.real() * b.imag() + (long double)a.imag() * b.real();
      }
      Cm[m + M * n] = C<F>((F)re, (F)im);
    }
}

// Loop_Mugiq::computeCoarseLoop, lib/loop_mugiq.cpp:455-509, in the reference's own schedule:
// for each entry (-1 = ultra-local) zero the entry's slots, then for each eigenvector copy it, displace it
// hop by hop (Displace::doVectorDisplacement, lib/displace.cpp:55-67) and contract where start<=k<=stop.
template <typename F>
void compute_loop(C<F> *dataPos, const C<F> *const *evecs, const double *sigma, int nEv, const C<F> *gauge, int nEntries,
                  const int *dir, const int *sign, const int *start, const int *stop, const Geom &g) {
  const size_t perLoop = (size_t)N_GAMMA * g.volume;
  std::vector<C<F>> R((size_t)g.volume * 12), aux((size_t)g.volume * 12);
  // nLoopOffset, include/loop_mugiq.h:244-247
  std::vector<int> offset(nEntries > 0 ? nEntries : 1);
  int osum = 1;
  for (int id = 0; id < nEntries; id++) {
    offset[id] = osum;
    osum += stop[id] - start[id] + 1;
  }
  for (int id = -1; id < nEntries; id++) {
    const size_t bufOffset = (id < 0) ? 0 : perLoop * offset[id];
    const size_t nSlots = (id < 0) ? 1 : (size_t)(stop[id] - start[id] + 1);
    std::memset((void *)(dataPos + bufOffset), 0, sizeof(C<F>) * perLoop * nSlots);
    for (int n = 0; n < nEv; n++) {
      const F sg = (F)sigma[n];
      const C<F> *L = evecs[n];
      if (id >= 0) {
        std::memcpy((void *)R.data(), (const void *)L, sizeof(C<F>) * R.size());
        int dispCount = 0;
        for (int idisp = 1; idisp <= stop[id]; idisp++) {
          displace(aux.data(), R.data(), gauge, dir[id], sign[id], g);
          R.swap(aux);
          if (idisp >= start[id] && idisp <= stop[id]) {
            contract(dataPos + bufOffset + perLoop * dispCount, L, R.data(), sg, g);
            dispCount++;
          }
        }
      } else {
        contract(dataPos, L, L, sg, g);
      }
    }
  }
}

}  // namespace

#define ORC_EXPORT extern "C" __attribute__((visibility("default")))

ORC_EXPORT int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_EXPORT void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

ORC_EXPORT void orc_gamma_tables(double *row_value, int *column_index, double *map_sign, int *map_index) {
  for (int G = 0; G < 16; G++)
    for (int s = 0; s < 4; s++) {
      row_value[(G * 4 + s) * 2 + 0] = kRowValue[G][s][0];
      row_value[(G * 4 + s) * 2 + 1] = kRowValue[G][s][1];
      column_index[G * 4 + s] = kColumnIdx[G][s];
    }
  gamma_map(map_sign, map_index);
}

ORC_EXPORT void orc_get_coords(int *x, int cb, const int *L, int parity) { get_coords(x, cb, L, parity); }
ORC_EXPORT int orc_cb_index(const int *x, const int *L) { return link_index(x, L); }

#define ORC_INSTANTIATE(SUF, F)                                                                                     \
  ORC_EXPORT void orc_contract_##SUF(F *loop, const F *vL, const F *vR, double sigma, const int *L) {               \
    contract<F>((C<F> *)loop, (const C<F> *)vL, (const C<F> *)vR, (F)sigma, Geom(L));                               \
  }                                                                                                                 \
  ORC_EXPORT void orc_displace_##SUF(F *dst, const F *src, const F *gauge, int dir, int sign, const int *L) {       \
    displace<F>((C<F> *)dst, (const C<F> *)src, (const C<F> *)gauge, dir, sign, Geom(L));                           \
  }                                                                                                                 \
  ORC_EXPORT void orc_reorder_mapgamma_##SUF(F *out, const F *in, int nLoop, const int *L) {                        \
    reorder_mapgamma<F>((C<F> *)out, (const C<F> *)in, nLoop, Geom(L));                                             \
  }                                                                                                                 \
  ORC_EXPORT void orc_phase_matrix_##SUF(F *phase, const int *mom, int Nmom, int ftsign, const int *localL,         \
                                         const int *totalL, const int *commCoord) {                                 \
    phase_matrix<F>((C<F> *)phase, mom, Nmom, ftsign, localL, totalL, commCoord);                                   \
  }                                                                                                                 \
  ORC_EXPORT void orc_gemm_##SUF(F *Cm, const F *A, const F *B, long long M, int N, long long K) {                  \
    gemm<F>((C<F> *)Cm, (const C<F> *)A, (const C<F> *)B, M, N, K);                                                 \
  }                                                                                                                 \
  ORC_EXPORT void orc_compute_loop_##SUF(F *dataPos, const F *evecs, long long evec_stride, const double *sigma,    \
                                         int nEv, const F *gauge, int nEntries, const int *dir, const int *sign,    \
                                         const int *start, const int *stop, const int *L) {                         \
    std::vector<const C<F> *> ptr(nEv);                                                                             \
    for (int n = 0; n < nEv; n++) ptr[n] = (const C<F> *)(evecs + (size_t)n * evec_stride);                         \
    compute_loop<F>((C<F> *)dataPos, ptr.data(), sigma, nEv, (const C<F> *)gauge, nEntries, dir, sign, start, stop, \
                    Geom(L));                                                                                       \
  }

ORC_INSTANTIATE(f64, double)
ORC_INSTANTIATE(f32, float)
