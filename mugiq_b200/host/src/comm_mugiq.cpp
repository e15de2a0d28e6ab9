// comm_mugiq.cpp — NCCL-backed implementation of comm_mugiq.h.
#include "comm_mugiq.h"

#include <cuda_runtime.h>
#include <nccl.h>
#include <unistd.h>

#include <cstring>
#include <fstream>

#include "host_util.h"

struct MugiqComm {
  int rank, size;
  ncclComm_t nccl;
  cudaStream_t stream;
  float *flag_d = nullptr;  // operand of the stream barrier
};

#define NCCL_CHECK(expr)                                                             \
  do {                                                                               \
    ncclResult_t r_ = (expr);                                                        \
    if (r_ != ncclSuccess) errorQuda("%s failed: %s", #expr, ncclGetErrorString(r_)); \
  } while (0)

MugiqComm *mugiqCommInit(int rank, int size, int device, const char *id_file) {
  if (size < 1 || rank < 0 || rank >= size) errorQuda("mugiqCommInit: bad rank/size %d/%d", rank, size);
  HOST_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  if (rank == 0) {
    NCCL_CHECK(ncclGetUniqueId(&id));
    const std::string tmp = std::string(id_file) + ".tmp";
    {
      std::ofstream out(tmp, std::ios::binary);
      if (!out) errorQuda("mugiqCommInit: cannot write %s", tmp.c_str());
      out.write(reinterpret_cast<const char *>(&id), sizeof(id));
    }
    if (rename(tmp.c_str(), id_file) != 0) errorQuda("mugiqCommInit: cannot publish %s", id_file);
  } else {
    bool ok = false;
    for (int tries = 0; tries < 6000 && !ok; tries++) {  // up to 60 s
      std::ifstream in(id_file, std::ios::binary);
      if (in && in.read(reinterpret_cast<char *>(&id), sizeof(id))) ok = true;
      else usleep(10000);
    }
    if (!ok) errorQuda("mugiqCommInit: rank %d did not find the NCCL id file %s", rank, id_file);
  }
  MugiqComm *c = new MugiqComm;
  c->rank = rank;
  c->size = size;
  NCCL_CHECK(ncclCommInitRank(&c->nccl, size, id, rank));
  HOST_CUDA(cudaStreamCreate(&c->stream));
  return c;
}

void mugiqCommFinalize(MugiqComm *c) {
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  if (c->flag_d) cudaFree(c->flag_d);
  ncclCommDestroy(c->nccl);
  cudaStreamDestroy(c->stream);
  delete c;
}

int mugiqCommRank(const MugiqComm *c) { return c ? c->rank : 0; }
int mugiqCommSize(const MugiqComm *c) { return c ? c->size : 1; }

void mugiqCommAllReduceSum(MugiqComm *c, void *buf_d, size_t count, QudaPrecision prec) {
  if (!c || c->size == 1) return;
  HOST_CUDA(cudaDeviceSynchronize());  // the loop kernels run on the default stream
  NCCL_CHECK(ncclAllReduce(buf_d, buf_d, count, prec == QUDA_DOUBLE_PRECISION ? ncclDouble : ncclFloat, ncclSum, c->nccl, c->stream));
  HOST_CUDA(cudaStreamSynchronize(c->stream));
}

void mugiqCommAllGather(MugiqComm *c, const void *send_d, void *recv_d, size_t bytes) {
  if (!c || c->size == 1) {
    HOST_CUDA(cudaMemcpy(recv_d, send_d, bytes, cudaMemcpyDeviceToDevice));
    return;
  }
  HOST_CUDA(cudaDeviceSynchronize());
  NCCL_CHECK(ncclAllGather(send_d, recv_d, bytes, ncclChar, c->nccl, c->stream));
  HOST_CUDA(cudaStreamSynchronize(c->stream));
}

void mugiqCommStreamBarrier(MugiqComm *c, void *stream) {
  if (!c || c->size == 1) return;
  if (!c->flag_d) {
    HOST_CUDA(cudaMalloc((void **)&c->flag_d, sizeof(float)));
    HOST_CUDA(cudaMemset(c->flag_d, 0, sizeof(float)));
  }
  NCCL_CHECK(ncclAllReduce(c->flag_d, c->flag_d, 1, ncclFloat, ncclSum, c->nccl, (cudaStream_t)stream));
}

void mugiqCommShard(int nEv, int rank, int size, int *lo, int *hi) {
  const int base = nEv / size, rem = nEv % size;
  *lo = rank * base + (rank < rem ? rank : rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}

static bool g_loop_tsplit = false;
void setLoopTSplit(bool on) { g_loop_tsplit = on; }
bool getLoopTSplit() { return g_loop_tsplit; }

static MugiqComm *g_loop_comm = nullptr;
void setLoopComm(MugiqComm *comm) { g_loop_comm = comm; }
MugiqComm *getLoopComm() { return g_loop_comm; }
