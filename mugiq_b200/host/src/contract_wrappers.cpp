// contract_wrappers.cpp — the free "wrapper" functions Loop_Mugiq and Displace call
// (/root/reference/lib/contract_wrappers.cu, forward-declared in include/loop_mugiq.h:280-311 and
// include/displace.h:109-111), each a thin caller of one C-ABI entry point.  Like the reference they are
// synchronous (cudaDeviceSynchronize after the launch, lib/contract_wrappers.cu:71,110,151,192) and abort through
// errorQuda on failure; the fused path used by computeCoarseLoop does not go through them.
#include "host_util.h"
#include "loop_mugiq.h"

namespace {
// canonical (site-major) view of a field: the field itself, or a converted temporary
struct SiteView {
  const void *ptr = nullptr;
  void *tmp = nullptr;
  SiteView(quda::ColorSpinorField *f, const mugiq_b200_geom_t &geom) {
    if (f->FieldOrder() == QUDA_SPACE_SPIN_COLOR_FIELD_ORDER) {
      ptr = f->V();
    } else {
      HOST_CUDA(cudaMalloc(&tmp, f->Bytes()));
      MUGIQ_CHECK(mugiq_b200_ingest_spinor(tmp, f->V(), abi_order(f->FieldOrder()), &geom, nullptr));
      ptr = tmp;
    }
  }
  ~SiteView() {
    if (tmp) cudaFree(tmp);
  }
};
}  // namespace

// The gamma tables are compiled into the kernels: nothing to upload (lib/contract_wrappers.cu:6-47 copies them to
// __constant__ memory).  The calls stay so that reference-shaped callers keep compiling; they verify the tables.
template <typename Float> void copyGammaCoeffStructToSymbol() {
  double rv[16][4][2];
  int ci[16][4];
  MUGIQ_CHECK(mugiq_b200_gamma_tables(&rv[0][0][0], &ci[0][0], nullptr, nullptr));
  if (rv[0][0][0] != 1.0 || ci[15][3] != 3) errorQuda("copyGammaCoeffStructToSymbol: unexpected gamma tables");
  printfQuda("%s: Gamma coefficients are compiled into the kernels\n", __func__);
}
template <typename Float> void copyGammaMapStructToSymbol() {
  double sign[16];
  int index[16];
  MUGIQ_CHECK(mugiq_b200_gamma_tables(nullptr, nullptr, sign, index));
  if (index[0] != 15 || sign[3] != -1.0) errorQuda("copyGammaMapStructToSymbol: unexpected gamma map");
  printfQuda("%s: Gamma map is compiled into the kernels\n", __func__);
}

template <typename Float>
void createPhaseMatrixGPU(quda::complex<Float> *phaseMatrix_d, const int *momMatrix_h, long long locV3, int Nmom,
                          int FTSign, const int localL[], const int totalL[]) {
  if (locV3 != (long long)localL[0] * localL[1] * localL[2]) errorQuda("createPhaseMatrixGPU: locV3 does not match localL");
  const int commCoord[4] = {quda::comm_coord(0), quda::comm_coord(1), quda::comm_coord(2), quda::comm_coord(3)};
  MUGIQ_CHECK(mugiq_b200_phase_matrix(phaseMatrix_d, momMatrix_h, Nmom, FTSign, localL, totalL, commCoord,
                                      (int)precision_of<Float>(), nullptr));
  HOST_CUDA(cudaDeviceSynchronize());
}

template <typename Float, QudaFieldOrder fieldOrder>
void performLoopContraction(quda::complex<Float> *loopData_d, quda::ColorSpinorField *eVecL, quda::ColorSpinorField *eVecR,
                            Float sigma) {
  if (eVecL->SiteSubset() != QUDA_FULL_SITE_SUBSET || eVecR->SiteSubset() != QUDA_FULL_SITE_SUBSET)
    errorQuda("%s: This function supports only Full Site Subset spinors!", __func__);
  const mugiq_b200_geom_t geom = make_geom(eVecL->X(), precision_of<Float>());
  if (eVecL->FieldOrder() != QUDA_SPACE_SPIN_COLOR_FIELD_ORDER && eVecL->FieldOrder() == eVecR->FieldOrder()) {
    // QUDA-native FLOAT2 / FLOAT4 fields are contracted in place: no conversion, no scratch field
    const void *l = eVecL->V(), *r = eVecR->V();
    const double s = (double)sigma;
    MUGIQ_CHECK(mugiq_b200_contract_native(loopData_d, &l, eVecL == eVecR ? nullptr : &r, &s, 1, abi_order(eVecL->FieldOrder()), 1,
                                           &geom, nullptr));
  } else {
    SiteView L(eVecL, geom), R(eVecR, geom);
    MUGIQ_CHECK(mugiq_b200_contract(loopData_d, L.ptr, eVecL == eVecR ? L.ptr : R.ptr, (double)sigma, &geom, nullptr));
  }
  HOST_CUDA(cudaDeviceSynchronize());
}

template <typename Float>
void convertIdxOrder_mapGamma(quda::complex<Float> *dataPosMP_d, const quda::complex<Float> *dataPos_d, int nData,
                              int nLoop, int nParity, int volumeCB, const int localL[]) {
  if (nParity != 2) errorQuda("%s: This function supports only Full Site Subset spinors!", __func__);
  if ((long long)volumeCB * 2 != (long long)localL[0] * localL[1] * localL[2] * localL[3])
    errorQuda("%s: volumeCB does not match localL", __func__);
  const mugiq_b200_geom_t geom = make_geom(localL, precision_of<Float>());
  MUGIQ_CHECK(mugiq_b200_reorder_mapgamma(dataPosMP_d, dataPos_d, nData, nLoop, &geom, nullptr));
  HOST_CUDA(cudaDeviceSynchronize());
}

template <typename Float, QudaFieldOrder order>
void performCovariantDisplacementVector(quda::ColorSpinorField *dst, quda::ColorSpinorField *src,
                                        quda::cudaGaugeField *gauge, DisplaceDir dispDir, DisplaceSign dispSign) {
  if (dst->SiteSubset() != QUDA_FULL_SITE_SUBSET || src->SiteSubset() != QUDA_FULL_SITE_SUBSET)
    errorQuda("%s: This function supports only Full Site Subset spinors!", __func__);
  const mugiq_b200_geom_t geom = make_geom(src->X(), precision_of<Float>());
  if (src->FieldOrder() != QUDA_SPACE_SPIN_COLOR_FIELD_ORDER && src->FieldOrder() == dst->FieldOrder()) {
    // source and destination in the same QUDA-native order: displaced in place of the layout
    void *d = dst->V();
    const void *s = src->V();
    MUGIQ_CHECK(mugiq_b200_displace_native(&d, &s, 1, gauge->Gauge_p(), (int)dispDir, (int)dispSign, abi_order(src->FieldOrder()),
                                           &geom, nullptr));
  } else {
    if (dst->FieldOrder() != QUDA_SPACE_SPIN_COLOR_FIELD_ORDER)
      errorQuda("%s: a destination in a QUDA-native order needs a source in the same order", __func__);
    SiteView S(src, geom);
    MUGIQ_CHECK(mugiq_b200_displace(dst->V(), S.ptr, gauge->Gauge_p(), (int)dispDir, (int)dispSign, &geom, nullptr));
  }
  HOST_CUDA(cudaDeviceSynchronize());
}

#define INSTANTIATE_WRAPPERS(Float)                                                                                    \
  template void copyGammaCoeffStructToSymbol<Float>();                                                                 \
  template void copyGammaMapStructToSymbol<Float>();                                                                   \
  template void createPhaseMatrixGPU<Float>(quda::complex<Float> *, const int *, long long, int, int, const int[],     \
                                            const int[]);                                                              \
  template void convertIdxOrder_mapGamma<Float>(quda::complex<Float> *, const quda::complex<Float> *, int, int, int,   \
                                                int, const int[]);
#define INSTANTIATE_ORDERED(Float, order)                                                                               \
  template void performLoopContraction<Float, order>(quda::complex<Float> *, quda::ColorSpinorField *,                 \
                                                     quda::ColorSpinorField *, Float);                                 \
  template void performCovariantDisplacementVector<Float, order>(quda::ColorSpinorField *, quda::ColorSpinorField *,   \
                                                                 quda::cudaGaugeField *, DisplaceDir, DisplaceSign);
INSTANTIATE_WRAPPERS(double)
INSTANTIATE_WRAPPERS(float)
INSTANTIATE_ORDERED(double, QUDA_FLOAT2_FIELD_ORDER)
INSTANTIATE_ORDERED(double, QUDA_FLOAT4_FIELD_ORDER)
INSTANTIATE_ORDERED(double, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER)
INSTANTIATE_ORDERED(float, QUDA_FLOAT2_FIELD_ORDER)
INSTANTIATE_ORDERED(float, QUDA_FLOAT4_FIELD_ORDER)
INSTANTIATE_ORDERED(float, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER)
