// loop_mugiq.h — Loop_Mugiq<Float,fieldOrder> (/root/reference/include/loop_mugiq.h:12-136): buffers, the
// eigenvector x displacement loop, momentum projection and output of the disconnected-loop calculation.
// Same public surface (constructor, computeCoarseLoop, writeLoopsHDF5, destructor), same buffer index orders,
// same fatal-error behaviour; all arithmetic happens in libmugiq_b200.so behind the C-ABI of mugiq_b200.h.
#ifndef MUGIQ_B200_LOOP_MUGIQ_H
#define MUGIQ_B200_LOOP_MUGIQ_H
#include "displace.h"
#include "eigsolve_mugiq.h"
#include <vector>

#include "mugiq.h"

using namespace quda;

template <typename Float, QudaFieldOrder fieldOrder> class Loop_Mugiq {
  struct LoopComputeParam;
  LoopComputeParam *cPrm;
  Displace<Float, fieldOrder> *displace;
  Eigsolve_Mugiq *eigsolve;  // borrowed
  ColorSpinorField *refVec;  // borrowed: the field whose geometry is used throughout

  // single process: the "space" and "time" communicators of the reference (lib/loop_mugiq.cpp:62-88) each hold
  // this rank only, so MPI_Reduce / MPI_Gather / MPI_Bcast (:406-424) are copies
  int tCoord = 0;
  MuGiqBool IamTimeProcess = MUGIQ_BOOL_TRUE;
  MuGiqBool commsAreSet = MUGIQ_BOOL_FALSE;

  complex<Float> *dataPos_d = nullptr;      // device, x_eo + V4*(G + 16*iL)
  complex<Float> *dataPosMP_d = nullptr;    // device, t + Lt*(G' + 16*iL) + Lt*nData*v3
  complex<Float> *dataMom_d = nullptr;      // device, t + Lt*(G' + 16*iL) + Lt*nData*im
  complex<Float> *dataPos = nullptr;        // host copy of dataPos_d
  complex<Float> *dataMom_h = nullptr;      // host copy of dataMom_d
  complex<Float> *dataMom = nullptr;        // host, summed over "space" ranks
  complex<Float> *dataMom_bcast = nullptr;  // host, gathered over "time" ranks
  complex<Float> *phaseMatrix_d = nullptr;
  void *momWorkspace_d = nullptr;
  bool fusedMomProj = true;                 // stages 3+4 as one kernel on dataPos_d (no dataPosMP_d)
  void *evecStage_d = nullptr;  // site-major staging for QUDA-native eigenvectors
  // eigenvector shards, MUGIQ_B200_PEER_REDUCE=1: dataPos_d is an IPC-shareable allocation every rank maps, the overlapped
  // position-space sum moves its chunks with the copy engines (mugiq_b200_comm_attach_peers) instead of NCCL's kernels
  bool peerReduce = false;
  void *peerStage_d = nullptr;
  char peerPosHandle[64];  // CUDA IPC handle of dataPos_d
  std::vector<void *> peerOpened;
  void attachPeerReduce(mugiq_b200_loop_plan_t *plan);
  mugiq_b200_loop_feed_t *feed = nullptr;  // streamed eigenvectors (Eigsolve_Mugiq::setEvecProducer): device staging ring
  int feedBatch = 0;
  void *producerStream = nullptr;          // cudaStream_t the producer enqueues on
  const void *gaugeHost[4] = {nullptr, nullptr, nullptr, nullptr};  // borrowed (MugiqLoopParam::gauge): the T split cuts its slab

  const size_t SizeCplxFloat = sizeof(complex<Float>);
  long long nElemMomTotPerLoop, nElemMomLocPerLoop, nElemPosLocPerLoop;
  long long nElemMomTot, nElemMomLoc, nElemPosLoc, nElemPhMat;

  MuGiqBool MomProjDone;
  MuGiqBool writeDataPos, writeDataMom;
  std::string momSpaceFilename, posSpaceFilename;

  void setupComms();
  void printLoopComputeParams();
  void allocateDataMemory();
  void freeDataMemory();
  void copyGammaToConstMem();
  void createPhaseMatrix();
  void performMomentumProjection();
  void computeCoarseLoopTSplit();  // the eigenvectors are this rank's time slab (comm_dim(3) > 1 or setLoopTSplit)
  void writeLoopsHDF5_Mom();
  void writeLoopsHDF5_Pos();

public:
  Loop_Mugiq(MugiqLoopParam *loopParams_, Eigsolve_Mugiq *eigsolve_);
  ~Loop_Mugiq();
  void writeLoopsHDF5();
  void computeCoarseLoop();
  void resetRun();  // re-upload the host links and re-arm the projection for another computeCoarseLoop on this object

  // read access for callers and tests (the reference keeps the buffers private and only writes them to HDF5)
  const complex<Float> *hostDataPos() const { return dataPos; }
  const complex<Float> *hostDataMom() const { return dataMom_bcast; }
  int nLoop() const;
  long long numElemPos() const { return nElemPosLoc; }
  long long numElemMom() const { return nElemMomTot; }
};

template <typename Float> void copyGammaCoeffStructToSymbol();
template <typename Float> void copyGammaMapStructToSymbol();
template <typename Float>
void createPhaseMatrixGPU(complex<Float> *phaseMatrix_d, const int *momMatrix_h, long long locV3, int Nmom, int FTSign,
                          const int localL[], const int totalL[]);
template <typename Float, QudaFieldOrder fieldOrder>
void performLoopContraction(complex<Float> *loopData_d, ColorSpinorField *eVecL, ColorSpinorField *eVecR, Float sigma);
template <typename Float>
void convertIdxOrder_mapGamma(complex<Float> *dataPosMP_d, const complex<Float> *dataPos_d, int nData, int nLoop,
                              int nParity, int volumeCB, const int localL[]);
#endif
