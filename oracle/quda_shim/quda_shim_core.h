// quda_shim_core.h — TEST INFRASTRUCTURE (oracle/): the smallest restatement of QUDA's public interfaces that lets the
// reference's OWN kernel translation units
//     /root/reference/lib/contract_wrappers.cu, lib/mugiq_contract_kernels.cu,
//     lib/mugiq_displace_kernels.cu,            lib/mugiq_util_kernels.cu
// compile UNMODIFIED (oracle/Makefile target `ref`, output oracle/_ref/libmugiq_ref.so).  QUDA itself is not in this
// image and its version is unpinned by the reference (CMakeLists.txt:112-114), so everything in THIS file is an
// assumption about QUDA, restated from QUDA's public definitions of the develop branch of early 2020:
//   * complex<T>: (x, y) = (re, im), usual arithmetic;
//   * colorspinor::FieldOrderCB<Float,4,3,1,order>(parity, x_cb, s, c) -> complex&, native orders
//       FLOAT2: parity*12*VolumeCB + (s*3 + c)*VolumeCB + x_cb
//       FLOAT4: parity*12*VolumeCB + (((s*3 + c)/2)*VolumeCB + x_cb)*2 + (s*3 + c)%2          (no padding);
//   * gauge accessor U(dir, x_cb, parity) -> 3x3 link matrix with element (row, col); the device storage order is
//     internal to QUDA, here it is [dir][parity][x_cb][row][col] (the host QDP order the reference uploads);
//   * Matrix: conj(M) is the Hermitian conjugate, (M * v)(s, c) = sum_c' M(c, c') v(s, c'), ColorSpinor::data[s*3 + c];
//   * getCoords / linkIndex / linkIndexP1 / linkIndexM1 / linkIndexShift: even/odd site order, checkerboard index =
//     lexicographic >> 1, periodic wrap; single process: comm_dim_partitioned = 0, comm_coord = 0, no ghost zones.
// What this buys: parity against the reference's own arithmetic (gamma tables and how they are assembled, colour trace
// and projection, which neighbour / link / parity / dagger a displacement picks, the reorder index and sign map, the
// phase formula) with the reference's own launch geometry; what it cannot pin is QUDA's side of the list above.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

// ---- enum_quda.h / quda.h (subset) -------------------------------------------------------------------------------
typedef enum QudaFieldOrder_s {
  QUDA_FLOAT_FIELD_ORDER = 1,
  QUDA_FLOAT2_FIELD_ORDER = 2,
  QUDA_FLOAT4_FIELD_ORDER = 4,
  QUDA_INVALID_FIELD_ORDER = -1
} QudaFieldOrder;
typedef enum QudaReconstructType_s { QUDA_RECONSTRUCT_NO = 18, QUDA_RECONSTRUCT_INVALID = -1 } QudaReconstructType;
typedef enum QudaGhostExchange_s {
  QUDA_GHOST_EXCHANGE_NO,
  QUDA_GHOST_EXCHANGE_PAD,
  QUDA_GHOST_EXCHANGE_EXTENDED
} QudaGhostExchange;
typedef enum QudaParity_s { QUDA_EVEN_PARITY = 0, QUDA_ODD_PARITY, QUDA_INVALID_PARITY } QudaParity;
typedef enum QudaSiteSubset_s { QUDA_PARITY_SITE_SUBSET = 1, QUDA_FULL_SITE_SUBSET = 2 } QudaSiteSubset;
typedef enum QudaPrecision_s { QUDA_SINGLE_PRECISION = 4, QUDA_DOUBLE_PRECISION = 8 } QudaPrecision;
typedef struct QudaGaugeParam_s QudaGaugeParam;          // only named in declarations of include/mugiq.h
typedef struct QudaEigParam_s QudaEigParam;
typedef struct QudaMultigridParam_s QudaMultigridParam;

// ---- util_quda.h (subset) -------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define errorQuda(...)        \
  do {                        \
    printf("ERROR: ");        \
    printf(__VA_ARGS__);      \
    printf("\n");             \
    __trap();                 \
  } while (0)
#else
#define errorQuda(...)                                      \
  do {                                                      \
    fprintf(stderr, "ERROR: ");                             \
    fprintf(stderr, __VA_ARGS__);                           \
    fprintf(stderr, " (%s:%d)\n", __FILE__, __LINE__);      \
    exit(1);                                                \
  } while (0)
#endif
#define warningQuda(...) do { fprintf(stderr, "WARNING: "); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define printfQuda(...) do { printf(__VA_ARGS__); } while (0)
#define checkCudaError()                                                                   \
  do {                                                                                     \
    cudaError_t e_ = cudaGetLastError();                                                   \
    if (e_ != cudaSuccess) errorQuda("(CUDA) %s", cudaGetErrorString(e_));                 \
  } while (0)

inline int comm_dim_partitioned(int) { return 0; }
inline int comm_coord(int) { return 0; }

namespace quda {

// ---- complex ----------------------------------------------------------------------------------------------------------
template <typename T> struct alignas(2 * sizeof(T)) complex {
  T x, y;
  complex() = default;
  __host__ __device__ complex(T re, T im = T(0)) : x(re), y(im) {}
  __host__ __device__ T real() const { return x; }
  __host__ __device__ T imag() const { return y; }
  __host__ __device__ complex &operator+=(const complex &b) {
    x += b.x;
    y += b.y;
    return *this;
  }
  __host__ __device__ complex &operator=(T re) {
    x = re;
    y = T(0);
    return *this;
  }
};
template <typename T> __host__ __device__ inline complex<T> operator+(const complex<T> &a, const complex<T> &b) {
  return complex<T>(a.x + b.x, a.y + b.y);
}
template <typename T> __host__ __device__ inline complex<T> operator*(const complex<T> &a, const complex<T> &b) {
  return complex<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> __host__ __device__ inline complex<T> operator*(const T &a, const complex<T> &b) {
  return complex<T>(a * b.x, a * b.y);
}
template <typename T> __host__ __device__ inline complex<T> operator*(const complex<T> &a, const T &b) {
  return complex<T>(a.x * b, a.y * b);
}
template <typename T> __host__ __device__ inline complex<T> conj(const complex<T> &a) { return complex<T>(a.x, -a.y); }

// ---- fields (host-side handles around device memory) --------------------------------------------------------------------
class ColorSpinorField {
  void *v_;
  int x_[4];
  QudaFieldOrder order_;
  QudaPrecision prec_;

 public:
  ColorSpinorField(void *v, const int x[4], QudaFieldOrder order, QudaPrecision prec) : v_(v), order_(order), prec_(prec) {
    for (int i = 0; i < 4; i++) x_[i] = x[i];
  }
  void *V() const { return v_; }
  int X(int i) const { return x_[i]; }
  const int *X() const { return x_; }
  int Volume() const { return x_[0] * x_[1] * x_[2] * x_[3]; }
  int VolumeCB() const { return Volume() / 2; }
  int SiteSubset() const { return QUDA_FULL_SITE_SUBSET; }
  QudaFieldOrder FieldOrder() const { return order_; }
  QudaPrecision Precision() const { return prec_; }
  void exchangeGhost(QudaParity, int, int) const {}  // single process: nothing to exchange
};

class cudaGaugeField {
  void *g_;
  int x_[4], r_[4];
  QudaGhostExchange ghost_;

 public:
  cudaGaugeField(void *g, const int x[4], QudaGhostExchange ghost) : g_(g), ghost_(ghost) {
    for (int i = 0; i < 4; i++) {
      x_[i] = x[i];
      r_[i] = 0;  // exRng[d] = 2 * partitioned(d) = 0 (lib/displace.cpp:16)
    }
  }
  void *Gauge_p() const { return g_; }
  const int *X() const { return x_; }
  const int *R() const { return r_; }
  int Volume() const { return x_[0] * x_[1] * x_[2] * x_[3]; }
  int SiteSubset() const { return QUDA_FULL_SITE_SUBSET; }
  QudaGhostExchange GhostExchange() const { return ghost_; }
};

// ---- color_spinor.h / quda_matrix.h ---------------------------------------------------------------------------------------
template <typename Float, int Nc, int Ns> struct ColorSpinor {
  complex<Float> data[Nc * Ns];
};
template <typename T, int N> struct Matrix {
  T data[N * N];
  __host__ __device__ T &operator()(int i, int j) { return data[i * N + j]; }
  __host__ __device__ const T &operator()(int i, int j) const { return data[i * N + j]; }
};
template <typename T, int N> __host__ __device__ inline Matrix<T, N> conj(const Matrix<T, N> &a) {  // Hermitian conjugate
  Matrix<T, N> r;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) r(i, j) = conj(a(j, i));
  return r;
}
template <typename Float, int Nc, int Ns>
__host__ __device__ inline ColorSpinor<Float, Nc, Ns> operator*(const Matrix<complex<Float>, Nc> &A,
                                                                const ColorSpinor<Float, Nc, Ns> &x) {
  ColorSpinor<Float, Nc, Ns> y;
  for (int s = 0; s < Ns; s++)
    for (int i = 0; i < Nc; i++) {
      complex<Float> acc(Float(0), Float(0));
      for (int j = 0; j < Nc; j++) acc += A(i, j) * x.data[s * Nc + j];
      y.data[s * Nc + i] = acc;
    }
  return y;
}

// ---- color_spinor_field_order.h -------------------------------------------------------------------------------------------
namespace colorspinor {
template <typename Float, int nSpin, int nColor, int nVec, QudaFieldOrder order> struct FieldOrderCB {
  complex<Float> *v;
  int volumeCB;
  FieldOrderCB(const ColorSpinorField &f) : v(static_cast<complex<Float> *>(f.V())), volumeCB(f.VolumeCB()) {}
  __host__ __device__ inline complex<Float> &operator()(int parity, int x_cb, int s, int c, int n = 0) const {
    const int j = (s * nColor + c) * nVec + n;
    const long long off = (long long)parity * nSpin * nColor * nVec * volumeCB;
    if (order == QUDA_FLOAT2_FIELD_ORDER) return v[off + (long long)j * volumeCB + x_cb];
    return v[off + ((long long)(j / 2) * volumeCB + x_cb) * 2 + j % 2];  // QUDA_FLOAT4_FIELD_ORDER
  }
  // ghost zones exist only for partitioned dimensions (never reached in a single process)
  __host__ __device__ inline complex<Float> &Ghost(int, int, int, int, int, int, int = 0) const { return v[0]; }
};
}  // namespace colorspinor

// ---- gauge_field_order.h ----------------------------------------------------------------------------------------------------
template <typename Float> struct GaugeAccessorShim {
  const complex<Float> *u;
  int volumeCB;
  GaugeAccessorShim(const cudaGaugeField &g) : u(static_cast<const complex<Float> *>(g.Gauge_p())), volumeCB(g.Volume() / 2) {}
  __host__ __device__ inline Matrix<complex<Float>, 3> operator()(int dir, int x_cb, int parity) const {
    Matrix<complex<Float>, 3> m;
    const complex<Float> *p = u + (((long long)dir * 2 + parity) * volumeCB + x_cb) * 9;
    for (int i = 0; i < 9; i++) m.data[i] = p[i];
    return m;
  }
  __host__ __device__ inline Matrix<complex<Float>, 3> Ghost(int dir, int, int parity) const { return (*this)(dir, 0, parity); }
};
template <typename Float, QudaReconstructType recon> struct gauge_mapper {
  typedef GaugeAccessorShim<Float> type;
};

// ---- index_helper.cuh ---------------------------------------------------------------------------------------------------------
template <typename I> __host__ __device__ inline void getCoords(int x[], int cb_index, const I X, int parity) {
  const int za = cb_index / (X[0] / 2);
  const int zb = za / X[1];
  x[1] = za - zb * X[1];
  x[3] = zb / X[2];
  x[2] = zb - x[3] * X[2];
  const int x1odd = (x[1] + x[2] + x[3] + parity) & 1;
  x[0] = 2 * cb_index + x1odd - za * X[0];
}
template <typename I> __host__ __device__ inline int linkIndex(const int x[], const I X) {
  return (((x[3] * X[2] + x[2]) * X[1] + x[1]) * X[0] + x[0]) >> 1;
}
template <typename I> __host__ __device__ inline int linkIndexShift(const int x[], const int dx[], const I X) {
  int y[4];
  for (int i = 0; i < 4; i++) y[i] = (x[i] + dx[i] + X[i]) % X[i];
  return (((y[3] * X[2] + y[2]) * X[1] + y[1]) * X[0] + y[0]) >> 1;
}
template <typename I> __host__ __device__ inline int linkIndexP1(const int x[], const I X, int mu) {
  int y[4] = {x[0], x[1], x[2], x[3]};
  y[mu] = (y[mu] + 1) % X[mu];
  return (((y[3] * X[2] + y[2]) * X[1] + y[1]) * X[0] + y[0]) >> 1;
}
template <typename I> __host__ __device__ inline int linkIndexM1(const int x[], const I X, int mu) {
  int y[4] = {x[0], x[1], x[2], x[3]};
  y[mu] = (y[mu] - 1 + X[mu]) % X[mu];
  return (((y[3] * X[2] + y[2]) * X[1] + y[1]) * X[0] + y[0]) >> 1;
}
// index into a ghost zone: only evaluated for partitioned dimensions
template <int dir, typename I> __host__ __device__ inline int ghostFaceIndex(const int[], const I, int, int) { return 0; }

}  // namespace quda
