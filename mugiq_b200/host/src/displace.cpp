// displace.cpp — Displace<F,order> (/root/reference/lib/displace.cpp): gauge upload, direction state machine,
// single-hop displacement, and ownership of the fused path's LoopPlan.
#include "displace.h"

#include <typeinfo>

#include "host_util.h"

template <typename F, QudaFieldOrder order>
Displace<F, order>::Displace(MugiqLoopParam *loopParams_, ColorSpinorField *csf_, QudaPrecision coarsePrec_)
    : dispString(""),
      gaugePtr{loopParams_->gauge[0], loopParams_->gauge[1], loopParams_->gauge[2], loopParams_->gauge[3]},
      qGaugePrm(loopParams_->gauge_param),
      gaugeField(nullptr),
      auxDispVec(nullptr),
      siteVec(nullptr) {
  printfQuda("%s: Precision is %s\n", __func__, typeid(F) == typeid(float) ? "single" : "double");
  if (!qGaugePrm) errorQuda("%s: gauge_param is not set", __func__);
  if (coarsePrec_ != precision_of<F>())
    errorQuda("%s: eigenvector precision %d does not match the Displace template (%zu bytes)", __func__, (int)coarsePrec_,
              sizeof(F));
  geom = make_geom(csf_->X(), precision_of<F>());
  // single process: no dimension is partitioned, the extended-halo field of the reference has border 0
  // (lib/displace.cpp:16) and equals the plain device copy
  gaugeField = createCudaGaugeField();
  printfQuda("%s: Gauge field has NOT extended Halo exchange\n", __func__);

  ColorSpinorParam csParam;
  for (int i = 0; i < 4; i++) csParam.x[i] = csf_->X(i);
  csParam.precision = coarsePrec_;
  csParam.fieldOrder = QUDA_SPACE_SPIN_COLOR_FIELD_ORDER;
  auxDispVec = ColorSpinorField::Create(csParam);
  blas::zero(*auxDispVec);
  if (order != QUDA_SPACE_SPIN_COLOR_FIELD_ORDER) siteVec = ColorSpinorField::Create(csParam);
}

template <typename F, QudaFieldOrder order> Displace<F, order>::~Displace() {
  for (int i = 0; i < N_DIM_; i++) gaugePtr[i] = nullptr;
  if (plan) mugiq_b200_loop_plan_destroy(plan);
  if (gaugeField) delete gaugeField;
  if (auxDispVec) delete auxDispVec;
  if (siteVec) delete siteVec;
}

template <typename F, QudaFieldOrder order> cudaGaugeField *Displace<F, order>::createCudaGaugeField() {
  if ((qGaugePrm->cuda_prec == QUDA_SINGLE_PRECISION && typeid(F) != typeid(float)) ||
      (qGaugePrm->cuda_prec == QUDA_DOUBLE_PRECISION && typeid(F) != typeid(double)))
    errorQuda("%s: Incompatible precision settings between Displace template %zu and gauge field parameters %d\n", __func__,
              sizeof(F), static_cast<int>(qGaugePrm->cuda_prec));
  if (qGaugePrm->cpu_prec != qGaugePrm->cuda_prec)
    errorQuda("%s: host links must already have the device precision (cpu_prec %d, cuda_prec %d)", __func__,
              (int)qGaugePrm->cpu_prec, (int)qGaugePrm->cuda_prec);
  for (int i = 0; i < N_DIM_; i++) {
    if (!gaugePtr[i]) errorQuda("%s: host gauge pointer %d is NULL", __func__, i);
    if (qGaugePrm->X[i] != geom.L[i])
      errorQuda("%s: gauge extent %d in dimension %d does not match the eigenvectors (%d)", __func__, qGaugePrm->X[i], i,
                geom.L[i]);
  }
  cudaGaugeField *g = new cudaGaugeField(geom.L, qGaugePrm->cuda_prec);
  MUGIQ_CHECK(mugiq_b200_gauge_upload(g->Gauge_p(), gaugePtr, &geom, nullptr));
  return g;
}

template <typename F, QudaFieldOrder order> void Displace<F, order>::reloadGauge() {
  MUGIQ_CHECK(mugiq_b200_gauge_upload(gaugeField->Gauge_p(), gaugePtr, &geom, nullptr));
}

template <typename F, QudaFieldOrder order> void Displace<F, order>::createLoopPlan(const std::vector<mugiq_b200_disp_entry_t> &entries) {
  if (plan) {
    mugiq_b200_loop_plan_destroy(plan);
    plan = nullptr;
  }
  MUGIQ_CHECK(mugiq_b200_loop_plan_create(&plan, gaugeField->Gauge_p(), entries.data(), (int)entries.size(), &geom, nullptr));
}

template <typename F, QudaFieldOrder order> void Displace<F, order>::resetAuxDispVec(ColorSpinorField *fineEvec) {
  if (fineEvec->FieldOrder() == QUDA_SPACE_SPIN_COLOR_FIELD_ORDER)
    *auxDispVec = *fineEvec;
  else
    MUGIQ_CHECK(mugiq_b200_ingest_spinor(auxDispVec->V(), fineEvec->V(), abi_order(fineEvec->FieldOrder()), &geom, nullptr));
  printfQuda("%s: Reset of auxilliary displaced vector done\n", __func__);
}

// The reference copies aux -> displacedEvec and back (two full-field copies, lib/displace.cpp:47-52); the result is
// the same when only the first copy is made, which is what happens here (layout conversion included).
template <typename F, QudaFieldOrder order> void Displace<F, order>::swapAuxDispVec(ColorSpinorField *displacedEvec) {
  if (displacedEvec->FieldOrder() == QUDA_SPACE_SPIN_COLOR_FIELD_ORDER)
    *displacedEvec = *auxDispVec;
  else
    MUGIQ_CHECK(mugiq_b200_export_spinor(displacedEvec->V(), abi_order(displacedEvec->FieldOrder()), auxDispVec->V(), &geom,
                                         nullptr));
}

template <typename F, QudaFieldOrder order>
void Displace<F, order>::doVectorDisplacement(DisplaceType dispType, ColorSpinorField *displacedEvec, int idisp) {
  if (dispType == DISPLACE_TYPE_COVARIANT) {
    if (dispDir == DispDirNone || dispSign == DispSignNone)
      errorQuda("%s: Got invalid dispDir and/or dispSign.\n", __func__);
    performCovariantDisplacementVector<F, order>(auxDispVec, displacedEvec, gaugeField, dispDir, dispSign);
    swapAuxDispVec(displacedEvec);
    printfQuda("%s: Step-%02d of a Covariant displacement done\n", __func__, idisp);
  } else {
    errorQuda("Unsupported Displacement type %d", static_cast<int>(dispType));
  }
}

template <typename F, QudaFieldOrder order> DisplaceFlag Displace<F, order>::WhichDisplaceFlag() {
  DisplaceFlag dFlag = DispFlagNone;
  for (int i = 0; i < (int)DisplaceFlagArray.size(); i++)
    if (dispString == DisplaceFlagArray[i]) dFlag = static_cast<DisplaceFlag>(i);
  if (dFlag == DispFlagNone) errorQuda("%s: Cannot parse given displacement string = %s.\n", __func__, dispString.c_str());
  return dFlag;
}

template <typename F, QudaFieldOrder order> DisplaceDir Displace<F, order>::WhichDisplaceDir() {
  if (dispFlag < DispFlag_X || dispFlag > DispFlag_t)
    errorQuda("%s: Unsupported/unrecongized displacement string %s and/or flag %d.\n", __func__, dispString.c_str(),
              static_cast<int>(dispFlag));
  return static_cast<DisplaceDir>(static_cast<int>(dispFlag) / 2);
}

template <typename F, QudaFieldOrder order> DisplaceSign Displace<F, order>::WhichDisplaceSign() {
  if (dispFlag < DispFlag_X || dispFlag > DispFlag_t)
    errorQuda("%s: Unsupported/unrecongized displacement string %s and/or flag %d.\n", __func__, dispString.c_str(),
              static_cast<int>(dispFlag));
  return (static_cast<int>(dispFlag) & 1) ? DispSignMinus : DispSignPlus;
}

template <typename F, QudaFieldOrder order> void Displace<F, order>::setupDisplacement(std::string dStr) {
  dispString = dStr;
  dispFlag = WhichDisplaceFlag();
  dispDir = WhichDisplaceDir();
  dispSign = WhichDisplaceSign();
  if (dispDir >= DispDir_x && dispDir <= DispDir_t && (dispSign == DispSignMinus || dispSign == DispSignPlus)) {
    static const char *dirName[N_DIM_] = {"x", "y", "z", "t"};
    printfQuda("%s: Displacement(s) will take place in the %s%s direction\n\n", __func__, dispSign == DispSignPlus ? "+" : "-",
               dirName[dispDir]);
  } else {
    errorQuda("%s: Got invalid dispDir and/or dispSign.\n", __func__);
  }
}

template class Displace<float, QUDA_FLOAT2_FIELD_ORDER>;
template class Displace<float, QUDA_FLOAT4_FIELD_ORDER>;
template class Displace<float, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER>;
template class Displace<double, QUDA_FLOAT2_FIELD_ORDER>;
template class Displace<double, QUDA_FLOAT4_FIELD_ORDER>;
template class Displace<double, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER>;
