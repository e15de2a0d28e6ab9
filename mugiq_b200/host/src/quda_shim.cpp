// quda_shim.cpp — storage behind quda_shim.h: logging with QUDA's fatal-error convention and the two field classes.
#include "quda_shim.h"

#include <cuda_runtime.h>

#include <cstring>

namespace quda {

static void default_handler(const char *msg) {
  fprintf(stderr, "%s\n", msg);
  fflush(stderr);
  exit(1);
}
static ErrorHandler g_handler = default_handler;
static QudaVerbosity g_verbosity = QUDA_SUMMARIZE;

ErrorHandler setErrorHandler(ErrorHandler h) {
  ErrorHandler old = g_handler;
  g_handler = h ? h : default_handler;
  return old;
}
void setVerbosityQuda(QudaVerbosity v) { g_verbosity = v; }

void errorQuda_(const char *file, int line, const char *fmt, ...) {
  char body[1024], msg[1280];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(body, sizeof(body), fmt, ap);
  va_end(ap);
  snprintf(msg, sizeof(msg), "MuGiq ERROR: %s (%s:%d)", body, file, line);
  g_handler(msg);
  abort();  // a handler that returns does not make the error recoverable
}
void warningQuda_(const char *fmt, ...) {
  if (g_verbosity == QUDA_SILENT) return;
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "MuGiq WARNING: ");
  vfprintf(stderr, fmt, ap);
  va_end(ap);
}
void printfQuda_(const char *fmt, ...) {
  if (g_verbosity < QUDA_VERBOSE) return;
  va_list ap;
  va_start(ap, fmt);
  vprintf(fmt, ap);
  va_end(ap);
}

#define SHIM_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t e_ = (expr);                                                   \
    if (e_ != cudaSuccess) errorQuda("%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

ColorSpinorField::ColorSpinorField(const ColorSpinorParam &p)
    : prec_(p.precision), order_(p.fieldOrder), subset_(p.siteSubset), nSpin_(p.nSpin), nColor_(p.nColor), v_(p.v),
      owned_(p.v == nullptr) {
  for (int i = 0; i < 4; i++) x_[i] = p.x[i];
  if (Volume() == 0) errorQuda("ColorSpinorField: zero volume");
  if (owned_) SHIM_CUDA(cudaMalloc(&v_, Bytes()));
}
ColorSpinorField::~ColorSpinorField() {
  if (owned_ && v_) cudaFree(v_);
}
ColorSpinorField &ColorSpinorField::operator=(const ColorSpinorField &src) {
  if (&src == this) return *this;
  if (src.Bytes() != Bytes() || src.order_ != order_ || src.prec_ != prec_)
    errorQuda("ColorSpinorField copy: incompatible fields");
  SHIM_CUDA(cudaMemcpy(v_, src.v_, Bytes(), cudaMemcpyDeviceToDevice));
  return *this;
}

cudaGaugeField::cudaGaugeField(const int x[4], QudaPrecision prec) : prec_(prec), gauge_(nullptr) {
  for (int i = 0; i < 4; i++) x_[i] = x[i];
  SHIM_CUDA(cudaMalloc(&gauge_, Bytes()));
}
cudaGaugeField::~cudaGaugeField() {
  if (gauge_) cudaFree(gauge_);
}

namespace blas {
void zero(ColorSpinorField &a) { SHIM_CUDA(cudaMemset(a.V(), 0, a.Bytes())); }
}  // namespace blas

bool checkGauge(const QudaGaugeParam *p) {
  if (!p) return false;
  for (int i = 0; i < 4; i++)
    if (p->X[i] < 1) return false;
  if (p->X[0] & 1) return false;
  return p->cpu_prec == QUDA_SINGLE_PRECISION || p->cpu_prec == QUDA_DOUBLE_PRECISION;
}

}  // namespace quda
