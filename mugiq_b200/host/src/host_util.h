// host_util.h — glue shared by the host-side sources: status -> errorQuda translation, geometry and layout helpers.
#pragma once
#include <cuda_runtime.h>

#include "mugiq.h"
#include "mugiq_b200.h"

// a non-zero status of the C-ABI is what the reference reports through errorQuda / checkCudaError()
#define MUGIQ_CHECK(call)                                                    \
  do {                                                                       \
    if ((call) < 0) errorQuda("%s: %s", #call, mugiq_b200_last_error());      \
  } while (0)
#define HOST_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) errorQuda("%s failed: %s", #expr, cudaGetErrorString(e_));     \
  } while (0)

inline mugiq_b200_geom_t make_geom(const int X[4], QudaPrecision prec) {
  mugiq_b200_geom_t g;
  for (int i = 0; i < 4; i++) g.L[i] = X[i];
  g.precision = (int)prec;
  return g;
}
template <typename Float> constexpr QudaPrecision precision_of() {
  return sizeof(Float) == 8 ? QUDA_DOUBLE_PRECISION : QUDA_SINGLE_PRECISION;
}
inline int abi_order(QudaFieldOrder o) {
  switch (o) {
    case QUDA_FLOAT2_FIELD_ORDER: return MUGIQ_B200_ORDER_FLOAT2;
    case QUDA_FLOAT4_FIELD_ORDER: return MUGIQ_B200_ORDER_FLOAT4;
    case QUDA_SPACE_SPIN_COLOR_FIELD_ORDER: return MUGIQ_B200_ORDER_SITE;
    default: errorQuda("Unsupported field order %d", (int)o);
  }
  return -1;
}
