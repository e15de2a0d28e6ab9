// eigsolve_mugiq.h — what Loop_Mugiq reads from the reference's Eigsolve_Mugiq (/root/reference/include/
// eigsolve_mugiq.h; uses at lib/loop_mugiq.cpp:42-43,442,449,479-483): the eigenvector fields, sigma_n = sqrt(lambda_n)
// and the eigenvector count.  QUDA's eigensolver / multigrid stay external inputs of the hot path (BASELINE.json
// north_star), so this class only CARRIES eigenpairs computed elsewhere.
#ifndef MUGIQ_B200_EIGSOLVE_MUGIQ_H
#define MUGIQ_B200_EIGSOLVE_MUGIQ_H
#include <vector>

#include "mugiq_api.h"

struct MugiqEigParam {
  QudaEigParam *QudaEigParams;
  int nEv;
  explicit MugiqEigParam(QudaEigParam *p) : QudaEigParams(p), nEv(p ? p->nEv : 0) {}
};

class Eigsolve_Mugiq {
  template <typename Float, QudaFieldOrder fieldOrder> friend class Loop_Mugiq;

  MugiqEigParam *eigParams;
  std::vector<quda::ColorSpinorField *> eVecs;  // borrowed device fields
  std::vector<double> *eVals_sigma;             // singular values sigma_n (only defined for MdagM / MMdag solves)
  std::vector<quda::ColorSpinorField *> tmpCSF; // unused without multigrid
  bool useMGenv = false;
  bool computeCoarse = false;

public:
  // eigenpairs handed in from outside: `sigma` may be empty (M / Mdag solves), then Loop_Mugiq refuses to run
  Eigsolve_Mugiq(MugiqEigParam *eigParams_, const std::vector<quda::ColorSpinorField *> &evecs,
                 const std::vector<double> &sigma);
  ~Eigsolve_Mugiq();
  std::vector<quda::ColorSpinorField *> &getEvecs() { return eVecs; }
  std::vector<double> *getEvalsSigma() { return eVals_sigma; }
  MugiqEigParam *getEigParams() { return eigParams; }
  void printInfo();
};

// computeLoop<Float>(mgParams, eigParams, ...) takes its eigenpairs from the object registered here.
void setExternalEigsolve(Eigsolve_Mugiq *eigsolve);
Eigsolve_Mugiq *getExternalEigsolve();

#endif
