// eigsolve_mugiq.h — what Loop_Mugiq reads from the reference's Eigsolve_Mugiq (/root/reference/include/
// eigsolve_mugiq.h; uses at lib/loop_mugiq.cpp:42-43,442,449,479-483): the eigenvector fields, sigma_n = sqrt(lambda_n)
// and the eigenvector count.  QUDA's eigensolver / multigrid stay external inputs of the hot path (BASELINE.json
// north_star), so this class only CARRIES eigenpairs computed elsewhere.
#ifndef MUGIQ_B200_EIGSOLVE_MUGIQ_H
#define MUGIQ_B200_EIGSOLVE_MUGIQ_H
#include <vector>

#include "mugiq.h"

struct MugiqEigParam {
  QudaEigParam *QudaEigParams;
  int nEv;
  explicit MugiqEigParam(QudaEigParam *p) : QudaEigParams(p), nEv(p ? p->nEv : 0) {}
};

// Streamed eigenvectors: the stand-in for Loop_Mugiq::prolongateEvec (/root/reference/lib/loop_mugiq.cpp:276-319, called
// per eigenvector at :482).  When a producer is registered, fine eigenvector n is PRODUCED - by QUDA's transfer operators
// in the reference; by whatever the caller plugs in here - into a device field the loop hands out (fine lattice, the
// loop's precision and field order), on the given CUDA stream (cudaStream_t as void*), right before the loop kernels
// consume it; it must only ENQUEUE work on that stream.  eVecs then only has to hold one field (the geometry reference,
// lib/loop_mugiq.cpp:42-43); the set does not have to fit the GPU.
typedef void (*EvecProducer)(void *ctx, int n, quda::ColorSpinorField *fineEvec, void *stream);

class Eigsolve_Mugiq {
  template <typename Float, QudaFieldOrder fieldOrder> friend class Loop_Mugiq;

  MugiqEigParam *eigParams;
  std::vector<quda::ColorSpinorField *> eVecs;  // borrowed device fields
  std::vector<double> *eVals_sigma;             // singular values sigma_n (only defined for MdagM / MMdag solves)
  std::vector<quda::ColorSpinorField *> tmpCSF; // unused without multigrid
  bool useMGenv = false;
  bool computeCoarse = false;
  EvecProducer producer = nullptr;
  void *producerCtx = nullptr;
  int producerBatch = 16;  // eigenvectors per staging batch of the streamed feed

public:
  // eigenpairs handed in from outside: `sigma` may be empty (M / Mdag solves), then Loop_Mugiq refuses to run
  Eigsolve_Mugiq(MugiqEigParam *eigParams_, const std::vector<quda::ColorSpinorField *> &evecs,
                 const std::vector<double> &sigma);
  ~Eigsolve_Mugiq();
  std::vector<quda::ColorSpinorField *> &getEvecs() { return eVecs; }
  std::vector<double> *getEvalsSigma() { return eVals_sigma; }
  MugiqEigParam *getEigParams() { return eigParams; }
  void printInfo();
  void setEvecProducer(EvecProducer fn, void *ctx, int batch = 16) {
    producer = fn;
    producerCtx = ctx;
    producerBatch = batch;
  }
};

// computeLoop<Float>(mgParams, eigParams, ...) takes its eigenpairs from the object registered here.
void setExternalEigsolve(Eigsolve_Mugiq *eigsolve);
Eigsolve_Mugiq *getExternalEigsolve();

#endif
