// comm_mugiq.cpp — implementation of comm_mugiq.h on the library's own communicator (mugiq_b200_comm_*, NCCL bound at
// run time by libmugiq_b200.so): this file holds the rendezvous and the bookkeeping, no NCCL call of its own.
#include "comm_mugiq.h"

#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "host_util.h"
#include "mugiq_b200.h"

struct MugiqComm {
  int rank, size;
  mugiq_b200_comm_t *lib;
  cudaStream_t stream;
  float *flag_d = nullptr;  // operand of the stream barrier
};

#define LIB_CHECK(expr)                                                    \
  do {                                                                     \
    if ((expr) != MUGIQ_B200_OK) errorQuda("%s failed: %s", #expr, mugiq_b200_last_error()); \
  } while (0)

// The id file carries a header that ties it to ONE launch: magic + the launch token every rank was given
// (MUGIQ_COMM_TOKEN, or the session id of the launching shell).  A file left behind by an earlier run with the same
// path has another token and is ignored by the readers; rank 0 removes its file once every rank has joined.
namespace {
struct IdFile {
  char magic[8];
  uint64_t token;
  char id[MUGIQ_B200_COMM_ID_BYTES];
};
uint64_t launch_token() {
  if (const char *e = getenv("MUGIQ_COMM_TOKEN")) return strtoull(e, nullptr, 0);
  return (uint64_t)getsid(0);
}
}  // namespace

MugiqComm *mugiqCommInit(int rank, int size, int device, const char *id_file) {
  if (size < 1 || rank < 0 || rank >= size) errorQuda("mugiqCommInit: bad rank/size %d/%d", rank, size);
  HOST_CUDA(cudaSetDevice(device));
  IdFile f;
  memcpy(f.magic, "MUGIQID2", 8);
  f.token = launch_token();
  if (rank == 0) {
    LIB_CHECK(mugiq_b200_comm_unique_id(f.id));
    const std::string tmp = std::string(id_file) + ".tmp";
    {
      std::ofstream out(tmp, std::ios::binary);
      if (!out) errorQuda("mugiqCommInit: cannot write %s", tmp.c_str());
      out.write(reinterpret_cast<const char *>(&f), sizeof(f));
    }
    if (rename(tmp.c_str(), id_file) != 0) errorQuda("mugiqCommInit: cannot publish %s", id_file);
  } else {
    bool ok = false;
    for (int tries = 0; tries < 6000 && !ok; tries++) {  // up to 60 s
      IdFile g;
      std::ifstream in(id_file, std::ios::binary);
      if (in && in.read(reinterpret_cast<char *>(&g), sizeof(g)) && memcmp(g.magic, f.magic, 8) == 0 && g.token == f.token) {
        f = g;
        ok = true;
      } else {
        usleep(10000);
      }
    }
    if (!ok)
      errorQuda("mugiqCommInit: rank %d found no NCCL id file %s of this launch (token %llu)", rank, id_file,
                (unsigned long long)f.token);
  }
  MugiqComm *c = new MugiqComm;
  c->rank = rank;
  c->size = size;
  LIB_CHECK(mugiq_b200_comm_create(&c->lib, f.id, rank, size));  // returns when every rank has joined
  if (rank == 0) unlink(id_file);
  HOST_CUDA(cudaStreamCreate(&c->stream));
  return c;
}

void mugiqCommFinalize(MugiqComm *c) {
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  if (c->flag_d) cudaFree(c->flag_d);
  mugiq_b200_comm_destroy(c->lib);
  cudaStreamDestroy(c->stream);
  delete c;
}

int mugiqCommRank(const MugiqComm *c) { return c ? c->rank : 0; }
int mugiqCommSize(const MugiqComm *c) { return c ? c->size : 1; }
mugiq_b200_comm_t *mugiqCommHandle(MugiqComm *c) { return c ? c->lib : nullptr; }

void mugiqCommAllReduceSum(MugiqComm *c, void *buf_d, size_t count, QudaPrecision prec) {
  if (!c || c->size == 1) return;
  HOST_CUDA(cudaDeviceSynchronize());  // the loop kernels run on the default stream
  LIB_CHECK(mugiq_b200_allreduce(buf_d, (long long)count, prec == QUDA_DOUBLE_PRECISION ? MUGIQ_B200_PREC_DOUBLE : MUGIQ_B200_PREC_SINGLE,
                                 c->lib, c->stream));
  HOST_CUDA(cudaStreamSynchronize(c->stream));
}

void mugiqCommAllGather(MugiqComm *c, const void *send_d, void *recv_d, size_t bytes) {
  if (!c || c->size == 1) {
    HOST_CUDA(cudaMemcpy(recv_d, send_d, bytes, cudaMemcpyDeviceToDevice));
    return;
  }
  HOST_CUDA(cudaDeviceSynchronize());
  LIB_CHECK(mugiq_b200_allgather(recv_d, send_d, (long long)bytes, c->lib, c->stream));
  HOST_CUDA(cudaStreamSynchronize(c->stream));
}

void mugiqCommStreamBarrier(MugiqComm *c, void *stream) {
  if (!c || c->size == 1) return;
  if (!c->flag_d) {
    HOST_CUDA(cudaMalloc((void **)&c->flag_d, sizeof(float)));
    HOST_CUDA(cudaMemset(c->flag_d, 0, sizeof(float)));
  }
  LIB_CHECK(mugiq_b200_allreduce(c->flag_d, 1, MUGIQ_B200_PREC_SINGLE, c->lib, stream));
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
