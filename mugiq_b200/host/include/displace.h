// displace.h — Displace<F,order> (/root/reference/include/displace.h:13-100): owns the device copy of the loop gauge
// field and the auxiliary displaced vector, parses "+x".."-t", applies one covariant hop at a time.
//
// On top of the reference's interface it owns the LoopPlan of the CUDA library (Wilson lines + launch schedule for
// the entry list), which is what Loop_Mugiq::computeCoarseLoop uses; doVectorDisplacement / swapAuxDispVec keep
// the reference's single-hop semantics for callers that drive the hops themselves.
#ifndef MUGIQ_B200_DISPLACE_H
#define MUGIQ_B200_DISPLACE_H
#include <string>
#include <vector>

#include "eigsolve_mugiq.h"
#include "mugiq.h"
#include "mugiq_b200.h"

using namespace quda;

template <typename F, QudaFieldOrder order> class Displace {
  template <typename Float, QudaFieldOrder fieldOrder> friend class Loop_Mugiq;

  const std::vector<std::string> DisplaceFlagArray{"+x", "-x", "+y", "-y", "+z", "-z", "+t", "-t"};
  const char *DisplaceTypeArray[N_DISPLACE_TYPES] = {"Covariant"};

  std::string dispString;
  DisplaceFlag dispFlag = DispFlagNone;
  DisplaceDir dispDir = DispDirNone;
  DisplaceSign dispSign = DispSignNone;

  void *gaugePtr[N_DIM_];         // borrowed host links (QDP order)
  QudaGaugeParam *qGaugePrm;      // borrowed
  cudaGaugeField *gaugeField;     // owned
  ColorSpinorField *auxDispVec;   // owned, canonical site-major order
  ColorSpinorField *siteVec;      // owned staging field when `order` is a QUDA native order
  mugiq_b200_geom_t geom;
  mugiq_b200_loop_plan_t *plan = nullptr;  // owned

  cudaGaugeField *createCudaGaugeField();
  void reloadGauge();  // H2D of the (borrowed) host links into the existing device field
  void setupDisplacement(std::string dStr);
  DisplaceFlag WhichDisplaceFlag();
  DisplaceDir WhichDisplaceDir();
  DisplaceSign WhichDisplaceSign();
  void resetAuxDispVec(ColorSpinorField *fineEvec);
  void doVectorDisplacement(DisplaceType dispType, ColorSpinorField *displacedEvec, int idisp);
  void swapAuxDispVec(ColorSpinorField *displacedEvec);
  // the fused path: plan for the entry list (dir, sign, start, stop per entry)
  void createLoopPlan(const std::vector<mugiq_b200_disp_entry_t> &entries);

public:
  Displace(MugiqLoopParam *loopParams_, ColorSpinorField *csf, QudaPrecision coarsePrec_);
  ~Displace();
};

template <typename Float, QudaFieldOrder order>
void performCovariantDisplacementVector(ColorSpinorField *dst, ColorSpinorField *src, cudaGaugeField *gauge,
                                        DisplaceDir dispDir, DisplaceSign dispSign);
#endif
