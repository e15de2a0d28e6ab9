// loop_mugiq.cpp — Loop_Mugiq<Float,fieldOrder> (/root/reference/lib/loop_mugiq.cpp).
//
// What is kept: constructor order (compute params -> comms -> buffers -> gamma tables -> phase matrix -> Displace,
// lib/loop_mugiq.cpp:7-59), buffer sizes and index orders (:102-158), the "always copy dataPos to the host" rule
// (:512), the call-once momentum projection (:325,433), the write-flag fix-ups (:669-693), fatal errors.
// What differs on purpose: the eigenvector x displacement loop nest (:455-509) is the LoopPlan of the CUDA library
// (one fused kernel per launch group and eigenvector batch instead of 4-6 synchronous launches and 3 field copies
// per eigenvector and hop); cuBLAS Zgemm/Cgemm (:364-377) is mugiq_b200_momproj; MPI and HDF5 are absent (single
// process; the loop file is the flat format of writeLoopsHDF5_Mom below).
#include "loop_mugiq.h"

#include <cstring>
#include <fstream>

#include "comm_mugiq.h"
#include "h5min.hpp"
#include "host_util.h"

template <typename Float, QudaFieldOrder fieldOrder> struct Loop_Mugiq<Float, fieldOrder>::LoopComputeParam {
  const int nG = N_GAMMA_;
  const int momDim = MOM_DIM_;
  int Nmom;
  LoopFTSign FTSign;
  std::vector<int> momMatrix;  // MOM_MATRIX_IDX(id, im) order
  MuGiqBool doMomProj, doNonLocal;
  int localL[N_DIM_], totalL[N_DIM_];
  int nParity, volumeCB;
  int locT = 0, totT = 0;
  long long locV4 = 1, locV3 = 1, totV3 = 1;
  LoopCalcType calcType;
  int nDispEntries = 0;
  std::vector<std::string> dispEntry, dispString;
  std::vector<int> dispStart, dispStop, nLoopPerEntry, nLoopOffset;
  int nLoop = 0, nData = 0;

  LoopComputeParam(MugiqLoopParam *prm, ColorSpinorField *x)
      : Nmom(prm->Nmom), FTSign(prm->FTSign), doMomProj(prm->doMomProj), doNonLocal(prm->doNonLocal),
        nParity(x->SiteSubset()), volumeCB((int)x->VolumeCB()), calcType(prm->calcType) {
    for (int i = 0; i < N_DIM_; i++) {
      localL[i] = x->X(i);
      totalL[i] = localL[i] * comm_dim(i);
      locV4 *= localL[i];
      if (i < N_DIM_ - 1) {
        locV3 *= localL[i];
        totV3 *= totalL[i];
      }
    }
    locT = localL[N_DIM_ - 1];
    totT = totalL[N_DIM_ - 1];
    if (doMomProj) {
      if (Nmom < 1 || (int)prm->momMatrix.size() != Nmom) errorQuda("Momentum matrix does not hold Nmom = %d momenta\n", Nmom);
      momMatrix.assign((size_t)Nmom * momDim, 0);
      for (int im = 0; im < Nmom; im++) {
        if ((int)prm->momMatrix[im].size() != momDim) errorQuda("Momentum %d does not have %d components\n", im, momDim);
        for (int id = 0; id < momDim; id++) momMatrix[MOM_MATRIX_IDX(id, im)] = prm->momMatrix[im][id];
      }
    }
    if (doNonLocal) {
      nDispEntries = (int)prm->disp_str.size();
      if (nDispEntries != (int)prm->disp_start.size() || nDispEntries != (int)prm->disp_stop.size())
        errorQuda("Displacement string length not compatible with displacement limits length\n");
      int offset = 1;  // slot 0 is the ultra-local loop
      for (int id = 0; id < nDispEntries; id++) {
        int start = prm->disp_start[id], stop = prm->disp_stop[id];
        if (start > stop) {
          warningQuda("Stop length is smaller than Start length for displacement %d. Will switch lengths!\n", id);
          std::swap(start, stop);
        }
        dispEntry.push_back(id < (int)prm->disp_entry.size() ? prm->disp_entry[id] : prm->disp_str[id]);
        dispString.push_back(prm->disp_str[id]);
        dispStart.push_back(start);
        dispStop.push_back(stop);
        nLoopPerEntry.push_back(stop - start + 1);
        nLoopOffset.push_back(offset);
        offset += stop - start + 1;
      }
      nLoop = offset;
    } else {
      nLoop = 1;
    }
    nData = nLoop * nG;
    printfQuda("%s: Loop compute parameters are set\n", __func__);
  }
};

template <typename Float, QudaFieldOrder fieldOrder>
Loop_Mugiq<Float, fieldOrder>::Loop_Mugiq(MugiqLoopParam *loopParams_, Eigsolve_Mugiq *eigsolve_)
    : cPrm(nullptr), displace(nullptr), eigsolve(eigsolve_), refVec(nullptr), nElemMomTotPerLoop(0), nElemMomLocPerLoop(0),
      nElemPosLocPerLoop(0), nElemMomTot(0), nElemMomLoc(0), nElemPosLoc(0), nElemPhMat(0), MomProjDone(MUGIQ_BOOL_FALSE),
      writeDataPos(loopParams_->writePosSpaceHDF5), writeDataMom(loopParams_->writeMomSpaceHDF5),
      momSpaceFilename(loopParams_->fname_mom_h5), posSpaceFilename(loopParams_->fname_pos_h5) {
  printfQuda("\n*************************************************\n");
  printfQuda("%s: Creating Loop computation environment\n", __func__);
  if (!eigsolve || eigsolve->eVecs.empty()) errorQuda("%s: no eigenvectors were given", __func__);
  if (eigsolve->useMGenv && eigsolve->computeCoarse)
    errorQuda("%s: coarse eigenvectors need QUDA's multigrid transfer, which is an external input of this build", __func__);
  refVec = eigsolve->eVecs[0];
  if (refVec->SiteSubset() != QUDA_FULL_SITE_SUBSET) errorQuda("%s: This class supports only Full Site Subset spinors!", __func__);
  if (refVec->FieldOrder() != fieldOrder) errorQuda("%s: eigenvector field order %d does not match the template", __func__, (int)refVec->FieldOrder());

  cPrm = new LoopComputeParam(loopParams_, refVec);
  for (int mu = 0; mu < 4; mu++) gaugeHost[mu] = loopParams_->gauge[mu];
  setupComms();
  allocateDataMemory();
  copyGammaToConstMem();
  if (cPrm->doMomProj) createPhaseMatrix();
  if (cPrm->doNonLocal) displace = new Displace<Float, fieldOrder>(loopParams_, refVec, refVec->Precision());
  printLoopComputeParams();
  printfQuda("*************************************************\n\n");
}

template <typename Float, QudaFieldOrder fieldOrder> Loop_Mugiq<Float, fieldOrder>::~Loop_Mugiq() {
  if (feed) mugiq_b200_loop_feed_destroy(feed);
  if (producerStream) cudaStreamDestroy((cudaStream_t)producerStream);
  freeDataMemory();
  if (displace) delete displace;
  if (cPrm) delete cPrm;
}

template <typename Float, QudaFieldOrder fieldOrder> int Loop_Mugiq<Float, fieldOrder>::nLoop() const { return cPrm->nLoop; }

// A second run on the same object (the reference builds one Loop_Mugiq per computeLoop call, lib/interface_mugiq.cpp:158-172):
// the host links are uploaded again (they may have changed) and the call-once guard of the projection is re-armed.
template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::resetRun() {
  if (displace) displace->reloadGauge();
  MomProjDone = MUGIQ_BOOL_FALSE;
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::setupComms() {
  // one rank: it is its own "space" and "time" communicator
  tCoord = comm_coord(3);
  IamTimeProcess = MUGIQ_BOOL_TRUE;
  commsAreSet = MUGIQ_BOOL_TRUE;
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::allocateDataMemory() {
  nElemPosLocPerLoop = cPrm->nG * cPrm->locV4;
  nElemMomLocPerLoop = (long long)cPrm->nG * cPrm->Nmom * cPrm->locT;
  nElemMomTotPerLoop = (long long)cPrm->nG * cPrm->Nmom * cPrm->totT;
  nElemPosLoc = nElemPosLocPerLoop * cPrm->nLoop;
  nElemMomLoc = nElemMomLocPerLoop * cPrm->nLoop;
  nElemMomTot = nElemMomTotPerLoop * cPrm->nLoop;
  nElemPhMat = (long long)cPrm->Nmom * cPrm->locV3;

  if (cPrm->doMomProj) {
    dataMom_h = static_cast<complex<Float> *>(calloc(nElemMomLoc, SizeCplxFloat));
    dataMom = static_cast<complex<Float> *>(calloc(nElemMomLoc, SizeCplxFloat));
    dataMom_bcast = static_cast<complex<Float> *>(calloc(nElemMomTot, SizeCplxFloat));
    if (!dataMom_h || !dataMom || !dataMom_bcast) errorQuda("%s: Could not allocate host momentum-space buffers\n", __func__);
    // stages 3+4 run as one kernel on dataPos_d (mugiq_b200_momproj_pos): no dataPosMP_d buffer.  MUGIQ_B200_UNFUSED_MOMPROJ=1
    // selects the reference's two-call form (convertIdxOrder_mapGamma, then the GEMM).
    fusedMomProj = !(getenv("MUGIQ_B200_UNFUSED_MOMPROJ") && getenv("MUGIQ_B200_UNFUSED_MOMPROJ")[0] == '1');
    HOST_CUDA(cudaMalloc((void **)&phaseMatrix_d, SizeCplxFloat * nElemPhMat));
    if (!fusedMomProj) HOST_CUDA(cudaMalloc((void **)&dataPosMP_d, SizeCplxFloat * nElemPosLoc));
    HOST_CUDA(cudaMalloc((void **)&dataMom_d, SizeCplxFloat * nElemMomLoc));
    const mugiq_b200_geom_t geom = make_geom(cPrm->localL, precision_of<Float>());
    const long long ws = fusedMomProj ? mugiq_b200_momproj_pos_workspace_bytes(&geom, cPrm->nLoop, cPrm->Nmom)
                                      : mugiq_b200_momproj_workspace_bytes((long long)cPrm->locT * cPrm->nData, cPrm->Nmom,
                                                                           cPrm->locV3, (int)precision_of<Float>());
    if (ws < 0) errorQuda("%s: %s", __func__, mugiq_b200_last_error());
    HOST_CUDA(cudaMalloc(&momWorkspace_d, (size_t)ws));
  }
  HOST_CUDA(cudaMallocHost((void **)&dataPos, SizeCplxFloat * nElemPosLoc));  // pinned: the D2H copy of :512 runs at PCIe speed
  const char *pr = getenv("MUGIQ_B200_PEER_REDUCE");
  peerReduce = pr && atoi(pr) != 0 && getLoopComm() && mugiqCommSize(getLoopComm()) > 1;
  if (peerReduce) {  // an allocation the other ranks can map (CUDA IPC)
    void *p = nullptr;
    MUGIQ_CHECK(mugiq_b200_peer_alloc(&p, (long long)(SizeCplxFloat * nElemPosLoc), peerPosHandle));
    dataPos_d = static_cast<complex<Float> *>(p);
  } else {
    HOST_CUDA(cudaMalloc((void **)&dataPos_d, SizeCplxFloat * nElemPosLoc));
  }
  HOST_CUDA(cudaMemset(dataPos_d, 0, SizeCplxFloat * nElemPosLoc));
  printfQuda("%s: Data buffers allocated\n", __func__);
}

// Peer transport of the overlapped position-space sum: the IPC handles of dataPos_d and of a staging area travel through the
// communicator's all-gather, every peer's pair is mapped and the tables go to the library (once per Loop_Mugiq).
template <typename Float, QudaFieldOrder fieldOrder>
void Loop_Mugiq<Float, fieldOrder>::attachPeerReduce(mugiq_b200_loop_plan_t *plan) {
  MugiqComm *comm = getLoopComm();
  const int world = mugiqCommSize(comm), rank = mugiqCommRank(comm);
  const long long stageBytes = mugiq_b200_comm_stage_bytes(plan, 8, world);
  if (stageBytes < 0) errorQuda("%s: %s", __func__, mugiq_b200_last_error());
  char mine[128];
  MUGIQ_CHECK(mugiq_b200_peer_alloc(&peerStage_d, stageBytes, mine + 64));
  memcpy(mine, peerPosHandle, 64);
  char *send_d = nullptr, *recv_d = nullptr;
  std::vector<char> all((size_t)128 * world);
  HOST_CUDA(cudaMalloc((void **)&send_d, 128));
  HOST_CUDA(cudaMalloc((void **)&recv_d, (size_t)128 * world));
  HOST_CUDA(cudaMemcpy(send_d, mine, 128, cudaMemcpyHostToDevice));
  mugiqCommAllGather(comm, send_d, recv_d, 128);
  HOST_CUDA(cudaMemcpy(all.data(), recv_d, all.size(), cudaMemcpyDeviceToHost));
  cudaFree(send_d);
  cudaFree(recv_d);
  std::vector<void *> pos(world), stage(world);
  for (int r = 0; r < world; r++) {
    if (r == rank) {
      pos[r] = dataPos_d;
      stage[r] = peerStage_d;
      continue;
    }
    MUGIQ_CHECK(mugiq_b200_peer_open(&pos[r], all.data() + (size_t)128 * r));
    MUGIQ_CHECK(mugiq_b200_peer_open(&stage[r], all.data() + (size_t)128 * r + 64));
    peerOpened.push_back(pos[r]);
    peerOpened.push_back(stage[r]);
  }
  MUGIQ_CHECK(mugiq_b200_comm_attach_peers(mugiqCommHandle(comm), pos.data(), stage.data(), stageBytes));
  printfQuda("%s: position-space sum over peer-mapped buffers (%d ranks, %.1f MB of staging per rank)\n", __func__, world, stageBytes / 1e6);
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::freeDataMemory() {
  auto dfree = [](auto *&p) {
    if (p) cudaFree((void *)p);
    p = nullptr;
  };
  auto hfree = [](auto *&p) {
    if (p) free((void *)p);
    p = nullptr;
  };
  hfree(dataMom_h);
  hfree(dataMom);
  hfree(dataMom_bcast);
  if (dataPos) cudaFreeHost(dataPos);
  dataPos = nullptr;
  if (peerReduce) {  // collective: nobody may still be copying into a buffer that goes away
    MugiqComm *comm = getLoopComm();
    cudaDeviceSynchronize();
    if (!peerOpened.empty()) mugiq_b200_comm_attach_peers(mugiqCommHandle(comm), nullptr, nullptr, 0);
    auto barrier = [&]() {  // stream-ordered one-element all-reduce, then wait for it
      mugiqCommStreamBarrier(comm, nullptr);
      cudaDeviceSynchronize();
    };
    barrier();
    for (void *p : peerOpened) mugiq_b200_peer_close(p);
    peerOpened.clear();
    barrier();
    if (peerStage_d) mugiq_b200_peer_free(peerStage_d);
    peerStage_d = nullptr;
    if (dataPos_d) mugiq_b200_peer_free(dataPos_d);
    dataPos_d = nullptr;
  }
  dfree(dataPos_d);
  dfree(dataPosMP_d);
  dfree(dataMom_d);
  dfree(phaseMatrix_d);
  dfree(momWorkspace_d);
  dfree(evecStage_d);
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::copyGammaToConstMem() {
  copyGammaCoeffStructToSymbol<Float>();
  copyGammaMapStructToSymbol<Float>();
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::createPhaseMatrix() {
  if (fusedMomProj) {  // same phases, stored in the even/odd run order the fused projection reads
    const int commCoord[4] = {comm_coord(0), comm_coord(1), comm_coord(2), comm_coord(3)};
    MUGIQ_CHECK(mugiq_b200_phase_matrix_eo(phaseMatrix_d, cPrm->momMatrix.data(), cPrm->Nmom, (int)cPrm->FTSign, cPrm->localL,
                                           cPrm->totalL, commCoord, (int)precision_of<Float>(), nullptr));
  } else {
    createPhaseMatrixGPU<Float>(phaseMatrix_d, cPrm->momMatrix.data(), cPrm->locV3, cPrm->Nmom, (int)cPrm->FTSign, cPrm->localL,
                                cPrm->totalL);
  }
  printfQuda("%s: Phase matrix created\n", __func__);
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::printLoopComputeParams() {
  printfQuda("Loop computation: %d loop(s) = ultra-local + %d displacement entr%s, %d momenta, momentum projection %s\n",
             cPrm->nLoop, cPrm->nDispEntries, cPrm->nDispEntries == 1 ? "y" : "ies", cPrm->Nmom, cPrm->doMomProj ? "on" : "off");
  for (int id = 0; id < cPrm->nDispEntries; id++)
    printfQuda("  entry %d: %s, lengths %d..%d, loops %d..%d\n", id, cPrm->dispString[id].c_str(), cPrm->dispStart[id],
               cPrm->dispStop[id], cPrm->nLoopOffset[id], cPrm->nLoopOffset[id] + cPrm->nLoopPerEntry[id] - 1);
  printfQuda("  local lattice %d %d %d %d, Fourier sign %d\n", cPrm->localL[0], cPrm->localL[1], cPrm->localL[2], cPrm->localL[3],
             (int)cPrm->FTSign);
}

// The eigenvector x displacement loop nest (lib/loop_mugiq.cpp:440-525).
template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::computeCoarseLoop() {
  const int nEv = eigsolve->eigParams->nEv;
  if (nEv < 1 || (!eigsolve->producer && nEv > (int)eigsolve->eVecs.size()))
    errorQuda("%s: nEv = %d but %zu eigenvectors were given", __func__, nEv, eigsolve->eVecs.size());
  if (!eigsolve->eVals_sigma || (int)eigsolve->eVals_sigma->size() < nEv)
    errorQuda("%s: singular values are only defined for MdagM / MMdag eigensolves and are required here", __func__);
  const QudaPrecision evecPrec = eigsolve->eVecs[0]->Precision();
  if (evecPrec != precision_of<Float>()) errorQuda("%s: Precision not supported!", __func__);
  if (getLoopTSplit()) {
    computeCoarseLoopTSplit();
    return;
  }
  const mugiq_b200_geom_t geom = make_geom(cPrm->localL, precision_of<Float>());

  // entries in the order given; directions parsed exactly as Displace does for the hop-by-hop interface
  std::vector<mugiq_b200_disp_entry_t> entries;
  for (int id = 0; id < cPrm->nDispEntries; id++) {
    displace->setupDisplacement(cPrm->dispString[id]);
    entries.push_back({(int)displace->dispDir, (int)displace->dispSign, cPrm->dispStart[id], cPrm->dispStop[id]});
  }
  mugiq_b200_loop_plan_t *plan = nullptr;
  if (displace) {
    displace->createLoopPlan(entries);
    plan = displace->plan;
  } else {
    MUGIQ_CHECK(mugiq_b200_loop_plan_create(&plan, nullptr, nullptr, 0, &geom, nullptr));
  }
  if (mugiq_b200_loop_plan_nloop(plan) != cPrm->nLoop) errorQuda("%s: plan holds %d loops, expected %d", __func__, mugiq_b200_loop_plan_nloop(plan), cPrm->nLoop);

  const bool sharded = getLoopComm() && mugiqCommSize(getLoopComm()) > 1;
  const size_t fieldBytes = eigsolve->eVecs[0]->Bytes();
  // FLOAT2 fields (the order computeLoop dispatches MG-coarse and double-precision eigenvectors to,
  // lib/interface_mugiq.cpp:226-235) are staged by the fused kernel itself; FLOAT4 fields are converted batch by batch
  const bool direct2 = fieldOrder == QUDA_FLOAT2_FIELD_ORDER && cPrm->volumeCB % 8 == 0;
  MUGIQ_CHECK(mugiq_b200_loop_plan_set_evec_order(plan, direct2 ? MUGIQ_B200_ORDER_FLOAT2 : MUGIQ_B200_ORDER_SITE));
  const bool native = fieldOrder != QUDA_SPACE_SPIN_COLOR_FIELD_ORDER && !direct2;
  if (eigsolve->producer) {
    // Streamed eigenvectors (the reference's prolongateEvec / field copy per eigenvector, lib/loop_mugiq.cpp:478-483,
    // batched): the library's feed hands out device staging batches, the producer fills batch b+1 on its stream while
    // the loop kernels consume batch b.
    const int batch = std::max(1, std::min({eigsolve->producerBatch, nEv, 256}));
    if (!feed || feedBatch != batch) {
      if (feed) mugiq_b200_loop_feed_destroy(feed);
      MUGIQ_CHECK(mugiq_b200_loop_feed_create(&feed, plan, dataPos_d, batch, 2,
                                              fieldOrder == QUDA_SPACE_SPIN_COLOR_FIELD_ORDER ? MUGIQ_B200_ORDER_SITE : abi_order(fieldOrder), 0,
                                              nullptr));
      feedBatch = batch;
    } else {
      MUGIQ_CHECK(mugiq_b200_loop_feed_set_plan(feed, plan, dataPos_d));
    }
    if (!producerStream) HOST_CUDA(cudaStreamCreateWithFlags((cudaStream_t *)&producerStream, cudaStreamNonBlocking));
    std::vector<void *> slot(batch);
    std::vector<double> sigma(batch);
    ColorSpinorParam cs;
    for (int i = 0; i < 4; i++) cs.x[i] = cPrm->localL[i];
    cs.precision = precision_of<Float>();
    cs.fieldOrder = fieldOrder;
    for (int n0 = 0; n0 < nEv; n0 += batch) {
      const int nb = std::min(batch, nEv - n0);
      MUGIQ_CHECK(mugiq_b200_loop_feed_acquire(feed, slot.data(), nb, producerStream));
      for (int i = 0; i < nb; i++) {
        cs.v = slot[i];  // wraps the staging field, not owned
        ColorSpinorField fine(cs);
        eigsolve->producer(eigsolve->producerCtx, n0 + i, &fine, producerStream);
        sigma[i] = (double)(Float)(*(eigsolve->eVals_sigma))[n0 + i];
      }
      MUGIQ_CHECK(mugiq_b200_loop_feed_commit(feed, sigma.data(), nb, producerStream));
      printfQuda("%s: Loop trace for eigenvectors %04d - %04d enqueued\n", __func__, n0, n0 + nb - 1);
    }
    long long fed = 0;
    MUGIQ_CHECK(mugiq_b200_loop_feed_finish(feed, &fed));
    if (fed != nEv) errorQuda("%s: the feed consumed %lld of %d eigenvectors", __func__, fed, nEv);
    if (sharded) {  // sum of the slots the plan computed, after the last batch
      std::vector<int> slots(cPrm->nLoop);
      const int ns = mugiq_b200_loop_plan_computed_slots(plan, slots.data(), cPrm->nLoop);
      if (ns < 1) errorQuda("%s: %s", __func__, mugiq_b200_last_error());
      MUGIQ_CHECK(mugiq_b200_allreduce_pos(dataPos_d, slots.data(), ns, 0, -1, &geom, mugiqCommHandle(getLoopComm()), nullptr));
    }
  } else {
  // eigenvectors are consumed in batches; QUDA-native orders are converted to the site-major layout batch by batch
  // Every batch costs one read-modify-write of the computed loop slots, so batches are as large as the kernel's
  // pointer table allows (256); QUDA-native fields need a site-major staging copy, which is capped at 4 GiB.
  int batch = 256;
  if (native) batch = (int)std::max<size_t>(8, std::min<size_t>(256, ((size_t)4 << 30) / fieldBytes));
  batch = std::min(batch, nEv);
  if (native && !evecStage_d) HOST_CUDA(cudaMalloc(&evecStage_d, fieldBytes * batch));
  std::vector<const void *> ptr(batch), src(batch);
  std::vector<void *> stage(batch);
  std::vector<double> sigma(batch);
  for (int n0 = 0; n0 < nEv; n0 += batch) {
    const int nb = std::min(batch, nEv - n0);
    for (int i = 0; i < nb; i++) {
      ColorSpinorField *v = eigsolve->eVecs[n0 + i];
      sigma[i] = (double)(Float)(*(eigsolve->eVals_sigma))[n0 + i];
      if (v->FieldOrder() != fieldOrder) errorQuda("%s: eigenvector %d has field order %d, expected %d", __func__, n0 + i, (int)v->FieldOrder(), (int)fieldOrder);
      src[i] = v->V();
      stage[i] = native ? static_cast<char *>(evecStage_d) + (size_t)i * fieldBytes : nullptr;
      ptr[i] = native ? stage[i] : v->V();
    }
    if (native)  // the whole batch in one launch
      MUGIQ_CHECK(mugiq_b200_ingest_spinor_batch(stage.data(), src.data(), nb, abi_order(fieldOrder), &geom, nullptr));
    // eigenvector shards (one process per GPU): the last batch's kernels run chunk by chunk in t, each chunk's cross-rank
    // sum over NVLink overlapping the next chunk's kernels (replaces the host-staged MPI collectives, lib/loop_mugiq.cpp:386-424)
    if (sharded && n0 + nb >= nEv) {
      if (peerReduce && peerOpened.empty()) attachPeerReduce(plan);
      MUGIQ_CHECK(mugiq_b200_loop_plan_accumulate_allreduce(plan, dataPos_d, ptr.data(), sigma.data(), nb, n0 > 0,
                                                            mugiqCommHandle(getLoopComm()), 8, nullptr));
    }
    else
      MUGIQ_CHECK(mugiq_b200_loop_plan_accumulate(plan, dataPos_d, ptr.data(), sigma.data(), nb, n0 > 0, nullptr));
    printfQuda("%s: Loop trace for eigenvectors %04d - %04d completed\n", __func__, n0, n0 + nb - 1);
  }
  }
  if (sharded) printfQuda("%s: Loop buffer summed over %d ranks\n", __func__, mugiqCommSize(getLoopComm()));
  // slots derived from computed ones (minus-direction partners, repeated entries) are linear in them: filled after the sum
  MUGIQ_CHECK(mugiq_b200_loop_plan_finalize(plan, dataPos_d, 0, nullptr));
  if (!displace) mugiq_b200_loop_plan_destroy(plan);

  // always copy the device position-space buffer to the host
  HOST_CUDA(cudaMemcpy(dataPos, dataPos_d, SizeCplxFloat * nElemPosLoc, cudaMemcpyDeviceToHost));
  HOST_CUDA(cudaDeviceSynchronize());
  printfQuda("%s: Loop trace completed\n", __func__);

  if (cPrm->doMomProj) {
    performMomentumProjection();
    printfQuda("%s: Momentum projection completed\n", __func__);
  }
}

// Lattice partitioned in t (the reference's comm_dim(3) > 1, SURVEY §8e secondary partitioning): the eigenvectors are this
// rank's time slab.  Replaces the per-hop exchangeGhostVec + ghost-zone reads (lib/contract_wrappers.cu:166-174,
// lib/mugiq_displace_kernels.cu:116-151) and the extended gauge field (lib/displace.cpp:104-134): the slabs are copied once
// into the extended layout [vector][parity][t = 0 .. Tl+2H)[V3/2][12] in an allocation the two time neighbours map through
// CUDA IPC; per eigenvector batch the boundary slices are written straight into the neighbours' halo slices over NVLink
// (copy engines), the kernels compute the interior only, and a minus-t loop derived from its plus partner fetches the
// partner's loop values below the interior from the neighbour once, after the eigenvector sum.
template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::computeCoarseLoopTSplit() {
  if (fieldOrder != QUDA_SPACE_SPIN_COLOR_FIELD_ORDER || sizeof(Float) != 8)
    errorQuda("%s: the lattice-T split of this build takes site-major double-precision eigenvectors", __func__);
  MugiqComm *comm = getLoopComm();
  const int world = comm_dim(3), rank = comm_coord(3);
  if (world != mugiqCommSize(comm) || (world > 1 && rank != mugiqCommRank(comm)))
    errorQuda("%s: %d time ranks announced but the communicator has %d", __func__, world, mugiqCommSize(comm));
  const int nEv = eigsolve->eigParams->nEv;
  const int Tl = cPrm->localL[3], T = cPrm->totalL[3];
  std::vector<mugiq_b200_disp_entry_t> entries;
  int tmax = 0;
  for (int id = 0; id < cPrm->nDispEntries; id++) {
    displace->setupDisplacement(cPrm->dispString[id]);
    entries.push_back({(int)displace->dispDir, (int)displace->dispSign, cPrm->dispStart[id], cPrm->dispStop[id]});
    if ((int)displace->dispDir == 3) tmax = std::max(tmax, cPrm->dispStop[id]);
  }
  const int H = tmax + (tmax & 1);  // even: a site keeps its parity in local, extended and global coordinates
  if (Tl % 2) errorQuda("%s: local T = %d must be even", __func__, Tl);
  if (H > Tl) errorQuda("%s: t-displacements of length %d exceed the local time extent %d", __func__, tmax, Tl);
  const int LtE = Tl + 2 * H;
  int Lext[4] = {cPrm->localL[0], cPrm->localL[1], cPrm->localL[2], LtE};
  const mugiq_b200_geom_t geomE = make_geom(Lext, precision_of<Float>());
  const size_t V3h = (size_t)cPrm->locV3 / 2, V4e = 2 * (size_t)LtE * V3h;
  const size_t S = 12 * SizeCplxFloat, U = 9 * SizeCplxFloat;

  // links of the extended slab, cut from the replicated global host field
  void *gaugeE_d = nullptr;
  if (cPrm->doNonLocal) {
    std::vector<char> slab(4 * V4e * U);
    const void *ptrs[4];
    for (int mu = 0; mu < 4; mu++) {
      if (!gaugeHost[mu]) errorQuda("%s: MugiqLoopParam::gauge[%d] is not set", __func__, mu);
      ptrs[mu] = slab.data() + (size_t)mu * V4e * U;
      for (int p = 0; p < 2; p++)
        for (int te = 0; te < LtE; te++) {
          const int tg = ((rank * Tl - H + te) % T + T) % T;
          memcpy(slab.data() + ((size_t)mu * V4e + ((size_t)p * LtE + te) * V3h) * U,
                 static_cast<const char *>(gaugeHost[mu]) + ((size_t)p * T + tg) * V3h * U, V3h * U);
        }
    }
    HOST_CUDA(cudaMalloc(&gaugeE_d, 4 * V4e * U));
    MUGIQ_CHECK(mugiq_b200_gauge_upload(gaugeE_d, ptrs, &geomE, nullptr));
  }

  // eigenvector slabs and the loop buffer in the extended layout, in allocations the neighbours can map
  void *slabs_d = nullptr, *posE_d = nullptr;
  char hSlabs[64], hPos[64];
  const size_t posBytes = (size_t)cPrm->nLoop * 16 * V4e * SizeCplxFloat;
  MUGIQ_CHECK(mugiq_b200_peer_alloc(&slabs_d, (long long)(nEv * V4e * S), hSlabs));
  MUGIQ_CHECK(mugiq_b200_peer_alloc(&posE_d, (long long)posBytes, hPos));
  HOST_CUDA(cudaMemset(posE_d, 0, posBytes));
  for (int n = 0; n < nEv; n++)
    for (int p = 0; p < 2; p++)
      HOST_CUDA(cudaMemcpyAsync(static_cast<char *>(slabs_d) + ((size_t)n * V4e + ((size_t)p * LtE + H) * V3h) * S,
                                static_cast<const char *>(eigsolve->eVecs[n]->V()) + (size_t)p * Tl * V3h * S, (size_t)Tl * V3h * S,
                                cudaMemcpyDeviceToDevice, nullptr));
  void *upSlabs = slabs_d, *dnSlabs = slabs_d, *upPos = posE_d;
  if (world > 1) {
    char *send_d = nullptr, *recv_d = nullptr;
    std::vector<char> all((size_t)128 * world);
    HOST_CUDA(cudaMalloc((void **)&send_d, 128));
    HOST_CUDA(cudaMalloc((void **)&recv_d, (size_t)128 * world));
    char mine[128];
    memcpy(mine, hSlabs, 64);
    memcpy(mine + 64, hPos, 64);
    HOST_CUDA(cudaMemcpy(send_d, mine, 128, cudaMemcpyHostToDevice));
    mugiqCommAllGather(comm, send_d, recv_d, 128);
    HOST_CUDA(cudaMemcpy(all.data(), recv_d, all.size(), cudaMemcpyDeviceToHost));
    cudaFree(send_d);
    cudaFree(recv_d);
    const int up = (rank + 1) % world, dn = (rank - 1 + world) % world;
    MUGIQ_CHECK(mugiq_b200_peer_open(&upSlabs, all.data() + (size_t)128 * up));
    MUGIQ_CHECK(mugiq_b200_peer_open(&upPos, all.data() + (size_t)128 * up + 64));
    if (dn == up)
      dnSlabs = upSlabs;
    else
      MUGIQ_CHECK(mugiq_b200_peer_open(&dnSlabs, all.data() + (size_t)128 * dn));
  }

  mugiq_b200_loop_plan_t *plan = nullptr;
  MUGIQ_CHECK(mugiq_b200_loop_plan_create(&plan, gaugeE_d, entries.empty() ? nullptr : entries.data(), (int)entries.size(), &geomE, nullptr));
  if (mugiq_b200_loop_plan_nloop(plan) != cPrm->nLoop) errorQuda("%s: plan holds %d loops, expected %d", __func__, mugiq_b200_loop_plan_nloop(plan), cPrm->nLoop);
  MUGIQ_CHECK(mugiq_b200_loop_plan_set_t_range(plan, H, H + Tl));  // the kernels compute the interior, the halos are read
  int lo = 0, up = 0, ll = 0;
  MUGIQ_CHECK(mugiq_b200_loop_plan_t_halo(plan, &lo, &up, &ll));

  cudaStream_t cs;
  HOST_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  HOST_CUDA(cudaDeviceSynchronize());  // the interiors are in place on every...
  mugiqCommStreamBarrier(comm, cs);    // ...rank before anybody writes halos next to them
  const int batch = std::min(nEv, 128);
  const int nbatch = (nEv + batch - 1) / batch;
  std::vector<cudaEvent_t> landed(nbatch);
  auto push = [&](int b) {  // halo slices of batch b -> the neighbours' slabs, then "everybody's pushes have landed"
    const int n0 = b * batch, nb = std::min(batch, nEv - n0);
    if (up) MUGIQ_CHECK(mugiq_b200_halo_push_t(dnSlabs, slabs_d, n0, nb, LtE, (long long)V3h, (int)S, H, H + Tl, up, 0, cs));
    if (lo) MUGIQ_CHECK(mugiq_b200_halo_push_t(upSlabs, slabs_d, n0, nb, LtE, (long long)V3h, (int)S, H + Tl - lo, H - lo, lo, 0, cs));
    mugiqCommStreamBarrier(comm, cs);
    HOST_CUDA(cudaEventCreateWithFlags(&landed[b], cudaEventDisableTiming));
    HOST_CUDA(cudaEventRecord(landed[b], cs));
  };
  push(0);
  std::vector<const void *> ptr(batch);
  std::vector<double> sigma(batch);
  for (int b = 0; b < nbatch; b++) {
    if (b + 1 < nbatch) push(b + 1);  // overlaps the kernels of batch b
    const int n0 = b * batch, nb = std::min(batch, nEv - n0);
    for (int i = 0; i < nb; i++) {
      ptr[i] = static_cast<const char *>(slabs_d) + (size_t)(n0 + i) * V4e * S;
      sigma[i] = (double)(Float)(*(eigsolve->eVals_sigma))[n0 + i];
    }
    HOST_CUDA(cudaStreamWaitEvent(nullptr, landed[b], 0));
    MUGIQ_CHECK(mugiq_b200_loop_plan_accumulate(plan, posE_d, ptr.data(), sigma.data(), nb, b > 0, nullptr));
    printfQuda("%s: Loop trace for eigenvectors %04d - %04d completed (time slab %d of %d)\n", __func__, n0, n0 + nb - 1, rank, world);
  }
  if (ll > 0) {
    // derived minus-t loops read their plus partner below the interior: the neighbour's top interior slices of every
    // plus-t slot, pushed with the same primitive (a loop slot is 16 "vectors" of 16-byte sites)
    HOST_CUDA(cudaDeviceSynchronize());
    int iL = 1;
    for (size_t id = 0; id < entries.size(); id++) {
      const int n = entries[id].stop - entries[id].start + 1;
      if (entries[id].dir == 3 && entries[id].sign == 1)
        for (int k = 0; k < n; k++)
          MUGIQ_CHECK(mugiq_b200_halo_push_t(upPos, posE_d, (iL + k) * 16, 16, LtE, (long long)V3h, (int)SizeCplxFloat, H + Tl - ll, H - ll, ll, 0, cs));
      iL += n;
    }
    mugiqCommStreamBarrier(comm, cs);
    HOST_CUDA(cudaStreamSynchronize(cs));
  }
  MUGIQ_CHECK(mugiq_b200_loop_plan_finalize(plan, posE_d, 0, nullptr));
  // the interior of the extended buffer is this rank's dataPos
  HOST_CUDA(cudaMemcpy2D(dataPos_d, (size_t)Tl * V3h * SizeCplxFloat, static_cast<const char *>(posE_d) + (size_t)H * V3h * SizeCplxFloat,
                         (size_t)LtE * V3h * SizeCplxFloat, (size_t)Tl * V3h * SizeCplxFloat, (size_t)cPrm->nLoop * 16 * 2, cudaMemcpyDeviceToDevice));
  HOST_CUDA(cudaMemcpy(dataPos, dataPos_d, SizeCplxFloat * nElemPosLoc, cudaMemcpyDeviceToHost));
  HOST_CUDA(cudaDeviceSynchronize());
  mugiqCommStreamBarrier(comm, cs);  // nobody unmaps while a neighbour may still write
  HOST_CUDA(cudaStreamSynchronize(cs));
  mugiq_b200_loop_plan_destroy(plan);
  for (cudaEvent_t e : landed) cudaEventDestroy(e);
  if (world > 1) {
    if (dnSlabs != upSlabs) MUGIQ_CHECK(mugiq_b200_peer_close(dnSlabs));
    MUGIQ_CHECK(mugiq_b200_peer_close(upSlabs));
    MUGIQ_CHECK(mugiq_b200_peer_close(upPos));
    mugiqCommStreamBarrier(comm, cs);  // everybody has unmapped before anybody frees
    HOST_CUDA(cudaStreamSynchronize(cs));
  }
  cudaStreamDestroy(cs);
  MUGIQ_CHECK(mugiq_b200_peer_free(slabs_d));
  MUGIQ_CHECK(mugiq_b200_peer_free(posE_d));
  if (gaugeE_d) cudaFree(gaugeE_d);
  printfQuda("%s: Loop trace completed (time slab %d of %d, halo %d slice(s) up, %d down, loop halo %d)\n", __func__, rank, world, up, lo, ll);

  if (cPrm->doMomProj) {
    performMomentumProjection();  // this rank's time-slices
    // MPI_Gather over COMM_TIME + MPI_Bcast (lib/loop_mugiq.cpp:420-424): every rank ends with all T time-slices
    if (world > 1) {
      complex<Float> *all_d = nullptr;
      HOST_CUDA(cudaMalloc((void **)&all_d, SizeCplxFloat * nElemMomLoc * world));
      mugiqCommAllGather(comm, dataMom_d, all_d, SizeCplxFloat * nElemMomLoc);
      std::vector<complex<Float>> all((size_t)nElemMomLoc * world);
      HOST_CUDA(cudaMemcpy(all.data(), all_d, SizeCplxFloat * all.size(), cudaMemcpyDeviceToHost));
      cudaFree(all_d);
      const size_t rows = (size_t)cPrm->Nmom * cPrm->nData;  // [im][idata] blocks of Tl (rank) -> T (gathered)
      for (int r = 0; r < world; r++)
        for (size_t k = 0; k < rows; k++)
          memcpy(dataMom_bcast + k * T + (size_t)r * Tl, all.data() + (size_t)r * nElemMomLoc + k * Tl, SizeCplxFloat * Tl);
    }
    printfQuda("%s: Momentum projection completed\n", __func__);
  }
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::performMomentumProjection() {
  if (MomProjDone) errorQuda("%s: Not supposed to be called more than once!!", __func__);
  if (!commsAreSet) setupComms();
  const long long locV3 = cPrm->locV3;
  const int locT = cPrm->locT, Nmom = cPrm->Nmom, nLoop = cPrm->nLoop, nData = cPrm->nData;
  if (nData != nLoop * N_GAMMA_) errorQuda("%s: This function assumes that nData = nLoop * NGamma\n", __func__);

  if (fusedMomProj) {
    const mugiq_b200_geom_t geom = make_geom(cPrm->localL, precision_of<Float>());
    MUGIQ_CHECK(mugiq_b200_momproj_pos(dataMom_d, dataPos_d, phaseMatrix_d, nLoop, Nmom, &geom, momWorkspace_d, nullptr));
  } else {
    // volume4d-inside-gamma-inside-nLoop -> time-inside-nData-inside-v3, with the Gamma -> g5*Gamma map
    convertIdxOrder_mapGamma<Float>(dataPosMP_d, dataPos_d, nData, nLoop, cPrm->nParity, cPrm->volumeCB, cPrm->localL);
    // dataMom(M x N) = dataPosMP(M x K) * phase(K x N), column-major
    const long long M = (long long)locT * nData, K = locV3;
    MUGIQ_CHECK(mugiq_b200_momproj(dataMom_d, dataPosMP_d, phaseMatrix_d, M, Nmom, K, (int)precision_of<Float>(), momWorkspace_d, nullptr));
  }
  HOST_CUDA(cudaMemcpy(dataMom_h, dataMom_d, SizeCplxFloat * nElemMomLoc, cudaMemcpyDeviceToHost));
  // MPI_Reduce over the "space" ranks and MPI_Gather + MPI_Bcast over the "time" ranks: one rank each
  memcpy(dataMom, dataMom_h, SizeCplxFloat * nElemMomLoc);
  memcpy(dataMom_bcast, dataMom, SizeCplxFloat * nElemMomLoc);
  MomProjDone = MUGIQ_BOOL_TRUE;
}

// HDF5 is not available in this build.  The same dataset tree the reference creates
// (/mom_%+d_%+d_%+d/disp_0 | disp_<dir>_<len>/<GammaName>/loop, shape [T][2], lib/loop_mugiq.cpp:582-633) is written
// as a flat file: a text index (one "dataset <path> <T> <offset>" line each) followed by the raw values.  Tags are
// not truncated (the reference's group2_tag[10] collides disp_+z_10 with disp_+z_1).  mugiq_b200/h5lite.py reads it.
template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::writeLoopsHDF5_Mom() {
  if (!MomProjDone) errorQuda("%s: momentum projection has not been performed", __func__);
  if (momSpaceFilename.empty()) errorQuda("%s: momentum-space file name is empty", __func__);
  const int totT = cPrm->totT, nData = cPrm->nData;
  std::vector<std::string> tags{"disp_0"};
  for (int id = 0; id < cPrm->nDispEntries; id++)
    for (int len = cPrm->dispStart[id]; len <= cPrm->dispStop[id]; len++) tags.push_back("disp_" + cPrm->dispString[id] + "_" + std::to_string(len));
  // a name ending in .h5 / .hdf5: a real HDF5 file with the reference's tree, written by the self-contained h5min.hpp
  auto endsWith = [&](const char *suf) {
    const size_t n = strlen(suf);
    return momSpaceFilename.size() >= n && momSpaceFilename.compare(momSpaceFilename.size() - n, n, suf) == 0;
  };
  if (endsWith(".h5") || endsWith(".hdf5")) {
    h5min::File h5;
    std::vector<Float> tmp((size_t)2 * totT);
    for (int im = 0; im < cPrm->Nmom; im++) {
      char momTag[64];
      snprintf(momTag, sizeof(momTag), "mom_%+d_%+d_%+d", cPrm->momMatrix[MOM_MATRIX_IDX(0, im)], cPrm->momMatrix[MOM_MATRIX_IDX(1, im)],
               cPrm->momMatrix[MOM_MATRIX_IDX(2, im)]);
      for (int iL = 0; iL < cPrm->nLoop; iL++)
        for (int ig = 0; ig < N_GAMMA_; ig++) {
          const complex<Float> *src = dataMom_bcast + (size_t)totT * ig + (size_t)totT * N_GAMMA_ * iL + (size_t)totT * nData * im;
          for (int t = 0; t < totT; t++) {
            tmp[2 * t] = src[t].real();
            tmp[2 * t + 1] = src[t].imag();
          }
          h5.addDataset("/" + std::string(momTag) + "/" + tags[iL] + "/" + GammaName()[ig] + "/loop", {(uint64_t)totT, 2}, (int)sizeof(Float),
                        tmp.data(), tmp.size() * sizeof(Float));
        }
    }
    if (!h5.write(momSpaceFilename)) errorQuda("%s: cannot open %s for writing", __func__, momSpaceFilename.c_str());
    printfQuda("%s: Momentum-space loops written to %s (HDF5)\n", __func__, momSpaceFilename.c_str());
    return;
  }
  std::string index = std::string("MUGIQ-B200 LOOPS v1 dtype ") + (sizeof(Float) == 8 ? "f64" : "f32") + "\n";
  std::vector<Float> payload;
  payload.reserve((size_t)2 * nElemMomTot);
  for (int im = 0; im < cPrm->Nmom; im++) {
    char momTag[64];
    snprintf(momTag, sizeof(momTag), "mom_%+d_%+d_%+d", cPrm->momMatrix[MOM_MATRIX_IDX(0, im)], cPrm->momMatrix[MOM_MATRIX_IDX(1, im)],
             cPrm->momMatrix[MOM_MATRIX_IDX(2, im)]);
    for (int iL = 0; iL < cPrm->nLoop; iL++)
      for (int ig = 0; ig < N_GAMMA_; ig++) {
        index += "dataset /" + std::string(momTag) + "/" + tags[iL] + "/" + GammaName()[ig] + "/loop " + std::to_string(totT) + " " +
                 std::to_string(payload.size() * sizeof(Float)) + "\n";
        const complex<Float> *src = dataMom_bcast + (size_t)totT * ig + (size_t)totT * N_GAMMA_ * iL + (size_t)totT * nData * im;
        for (int t = 0; t < totT; t++) {
          payload.push_back(src[t].real());
          payload.push_back(src[t].imag());
        }
      }
  }
  index += "end\n";
  std::ofstream out(momSpaceFilename, std::ios::binary);
  if (!out) errorQuda("%s: cannot open %s for writing", __func__, momSpaceFilename.c_str());
  out.write(index.data(), (std::streamsize)index.size());
  out.write(reinterpret_cast<const char *>(payload.data()), (std::streamsize)(payload.size() * sizeof(Float)));
  printfQuda("%s: Momentum-space loops written to %s\n", __func__, momSpaceFilename.c_str());
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::writeLoopsHDF5_Pos() {
  errorQuda("%s: Not supported yet!\n", __func__);  // as in the reference (lib/loop_mugiq.cpp:661-663)
}

template <typename Float, QudaFieldOrder fieldOrder> void Loop_Mugiq<Float, fieldOrder>::writeLoopsHDF5() {
  if (cPrm->doMomProj) {
    if (writeDataMom) {
      printfQuda("%s: Will write the momentum-space loop data\n", __func__);
    } else {
      warningQuda("%s: Performed momentum projection, but got writeDatMom = FALSE.\n", __func__);
      warningQuda("%s: Will proceed to write momentum-space loop data\n", __func__);
      writeDataMom = MUGIQ_BOOL_TRUE;
    }
    writeLoopsHDF5_Mom();
  } else if (!writeDataPos) {
    warningQuda("%s: Did not perform momentum projection, but got writeDatPos = FALSE.\n", __func__);
    warningQuda("%s: Will proceed to write position-space loop data\n", __func__);
    writeDataPos = MUGIQ_BOOL_TRUE;
  }
  if (writeDataPos) {
    printfQuda("%s: Will write the position-space loop data\n", __func__);
    writeLoopsHDF5_Pos();
  }
}

template class Loop_Mugiq<float, QUDA_FLOAT2_FIELD_ORDER>;
template class Loop_Mugiq<float, QUDA_FLOAT4_FIELD_ORDER>;
template class Loop_Mugiq<float, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER>;
template class Loop_Mugiq<double, QUDA_FLOAT2_FIELD_ORDER>;
template class Loop_Mugiq<double, QUDA_FLOAT4_FIELD_ORDER>;
template class Loop_Mugiq<double, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER>;
