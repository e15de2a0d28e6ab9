// quda_shim.h — the few QUDA types and helpers MuGiq's loop interface mentions, for building the host-side
// mirror (Loop_Mugiq / Displace / computeLoop) where QUDA itself is absent.
//
// The reference includes <quda.h>, <color_spinor_field.h>, <gauge_field.h> (e.g. /root/reference/include/mugiq.h:10,
// include/displace.h:4-8).  Only what the hot path touches is restated here: a ColorSpinorField is "a device
// pointer + lattice extents + precision + field order + site subset" (the accessors Loop_Mugiq / the wrappers call:
// V(), X(), VolumeCB(), Precision(), FieldOrder(), SiteSubset(); lib/contract_wrappers.cu:100,185,
// include/loop_mugiq.h:185-206).  With a real QUDA the same mirror compiles against QUDA's own headers instead
// (INTEGRATION.md): nothing below is used by the CUDA library.
#ifndef MUGIQ_B200_QUDA_SHIM_H
#define MUGIQ_B200_QUDA_SHIM_H

#include <complex>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <string>

// ---- enums (numerically equal to QUDA's enum_quda.h where the C-ABI relies on the value) -------------------------
typedef enum QudaPrecision_s {
  QUDA_QUARTER_PRECISION = 1,
  QUDA_HALF_PRECISION = 2,
  QUDA_SINGLE_PRECISION = 4,
  QUDA_DOUBLE_PRECISION = 8,
  QUDA_INVALID_PRECISION = -2147483647 - 1
} QudaPrecision;

typedef enum QudaFieldOrder_s {
  QUDA_FLOAT_FIELD_ORDER = 1,
  QUDA_FLOAT2_FIELD_ORDER = 2,   // native order of MG-coarse and double-precision fields
  QUDA_FLOAT4_FIELD_ORDER = 4,   // native order of single-precision fine fields
  QUDA_FLOAT8_FIELD_ORDER = 8,
  QUDA_SPACE_SPIN_COLOR_FIELD_ORDER,  // [parity][x_cb][spin][colour]: the CUDA library's canonical site-major order
  QUDA_INVALID_FIELD_ORDER = -2147483647 - 1
} QudaFieldOrder;

typedef enum QudaSiteSubset_s { QUDA_PARITY_SITE_SUBSET = 1, QUDA_FULL_SITE_SUBSET = 2 } QudaSiteSubset;
typedef enum QudaGaugeFieldOrder_s { QUDA_QDP_GAUGE_ORDER = 1 } QudaGaugeFieldOrder;
typedef enum QudaVerbosity_s { QUDA_SILENT, QUDA_SUMMARIZE, QUDA_VERBOSE, QUDA_DEBUG_VERBOSE } QudaVerbosity;

// ---- parameter structs (fields the loop path reads) ----------------------------------------------------------------
typedef struct QudaGaugeParam_s {
  int X[4];
  QudaPrecision cpu_prec;
  QudaPrecision cuda_prec;
  QudaGaugeFieldOrder gauge_order;
} QudaGaugeParam;

typedef struct QudaEigParam_s {
  int nEv;
} QudaEigParam;

typedef struct QudaMultigridParam_s {
  int n_level;
} QudaMultigridParam;

namespace quda {

template <typename T> using complex = std::complex<T>;

// ---- logging: errorQuda is fatal, as in QUDA (the reference has no return codes) -----------------------
typedef void (*ErrorHandler)(const char *msg);
ErrorHandler setErrorHandler(ErrorHandler h);  // tests install a throwing handler; default prints and exits
void errorQuda_(const char *file, int line, const char *fmt, ...);
void warningQuda_(const char *fmt, ...);
void printfQuda_(const char *fmt, ...);
void setVerbosityQuda(QudaVerbosity v);
#define errorQuda(...) ::quda::errorQuda_(__FILE__, __LINE__, __VA_ARGS__)
#define warningQuda(...) ::quda::warningQuda_(__VA_ARGS__)
#define printfQuda(...) ::quda::printfQuda_(__VA_ARGS__)

// The hot path shards eigenvectors, not the lattice; the one lattice partitioning it knows is the split in t
// (secondary partitioning), which the driver announces with setTimePartition(size, coord).
namespace detail {
inline int &time_ranks() {
  static int v = 1;
  return v;
}
inline int &time_coord() {
  static int v = 0;
  return v;
}
}  // namespace detail
inline void setTimePartition(int size, int coord) {
  detail::time_ranks() = size;
  detail::time_coord() = coord;
}
inline int comm_dim(int d) { return d == 3 ? detail::time_ranks() : 1; }
inline int comm_coord(int d) { return d == 3 ? detail::time_coord() : 0; }
inline int comm_rank() { return 0; }
inline int comm_size() { return 1; }

// ---- fields -----------------------------------------------------------------------------------------------------------
struct ColorSpinorParam {
  int x[4] = {0, 0, 0, 0};
  QudaPrecision precision = QUDA_DOUBLE_PRECISION;
  QudaFieldOrder fieldOrder = QUDA_SPACE_SPIN_COLOR_FIELD_ORDER;
  QudaSiteSubset siteSubset = QUDA_FULL_SITE_SUBSET;
  int nSpin = 4, nColor = 3;
  void *v = nullptr;  // wrap existing device memory (not owned) when non-null
};

class ColorSpinorField {
  int x_[4];
  QudaPrecision prec_;
  QudaFieldOrder order_;
  QudaSiteSubset subset_;
  int nSpin_, nColor_;
  void *v_;
  bool owned_;

public:
  explicit ColorSpinorField(const ColorSpinorParam &p);
  ~ColorSpinorField();
  ColorSpinorField(const ColorSpinorField &) = delete;
  ColorSpinorField &operator=(const ColorSpinorField &src);  // device copy; geometry and precision must match
  static ColorSpinorField *Create(const ColorSpinorParam &p) { return new ColorSpinorField(p); }
  void *V() { return v_; }
  const void *V() const { return v_; }
  const int *X() const { return x_; }
  int X(int d) const { return x_[d]; }
  size_t Volume() const { return (size_t)x_[0] * x_[1] * x_[2] * x_[3]; }
  size_t VolumeCB() const { return Volume() / 2; }
  size_t Length() const { return Volume() * nSpin_ * nColor_ * 2; }  // real numbers
  size_t Bytes() const { return Length() * (size_t)prec_; }
  QudaPrecision Precision() const { return prec_; }
  QudaFieldOrder FieldOrder() const { return order_; }
  QudaSiteSubset SiteSubset() const { return subset_; }
  int Nspin() const { return nSpin_; }
  int Ncolor() const { return nColor_; }
};

// Device gauge field in the order the CUDA library reads: [mu][parity][x_cb][row][col] complex.
class cudaGaugeField {
  int x_[4];
  QudaPrecision prec_;
  void *gauge_;

public:
  cudaGaugeField(const int x[4], QudaPrecision prec);
  ~cudaGaugeField();
  void *Gauge_p() { return gauge_; }
  const void *Gauge_p() const { return gauge_; }
  const int *X() const { return x_; }
  QudaPrecision Precision() const { return prec_; }
  size_t Bytes() const { return (size_t)4 * x_[0] * x_[1] * x_[2] * x_[3] * 18 * (size_t)prec_; }
};

namespace blas {
void zero(ColorSpinorField &a);
}

struct TimeProfile {
  std::string name;
  explicit TimeProfile(const std::string &n) : name(n) {}
};

}  // namespace quda
#endif
