// comm_mugiq.h — multi-GPU plumbing of the C++ host mirror: one process per GPU, eigenvectors sharded over the ranks,
// ONE NCCL all-reduce (sum) of the loop buffer over NVLink / NVSwitch.
//
// Replaces the reference's host-staged MPI_Reduce / MPI_Gather / MPI_Bcast of the momentum-space buffer
// (/root/reference/lib/loop_mugiq.cpp:406-424) and its two MPI_Comm_split communicators (:62-88): there is no spatial
// split to reduce over, the sum over eigenvector shards happens on the device buffer.  No MPI launcher exists in this
// image, so ranks rendezvous through a file that rank 0 writes the communicator id
// (mugiq_b200_comm_unique_id) to.
#ifndef MUGIQ_B200_COMM_MUGIQ_H
#define MUGIQ_B200_COMM_MUGIQ_H
#include <cstddef>

#include "mugiq.h"

struct MugiqComm;  // opaque
struct mugiq_b200_comm_s;

// The rendezvous file carries a launch token (environment MUGIQ_COMM_TOKEN, default: the session id of the launching
// shell, the same for every rank started from it): files left behind by other launches are ignored, and rank 0 removes
// its file once every rank has joined.

// Collective over all `size` ranks.  `device` is the CUDA device of this rank (cudaSetDevice is called).
MugiqComm *mugiqCommInit(int rank, int size, int device, const char *id_file);
void mugiqCommFinalize(MugiqComm *comm);
int mugiqCommRank(const MugiqComm *comm);
int mugiqCommSize(const MugiqComm *comm);
// the library communicator behind it (mugiq_b200_loop_plan_accumulate_allreduce, mugiq_b200_allreduce_pos)
mugiq_b200_comm_s *mugiqCommHandle(MugiqComm *comm);
// In-place sum of `count` real numbers of the given precision over all ranks (device buffer); returns when done.
void mugiqCommAllReduceSum(MugiqComm *comm, void *buf_d, size_t count, QudaPrecision prec);
// Lattice-T split (tsplit in loop_mugiq.cpp): every rank contributes `bytes` bytes (device buffers), rank-major result.
void mugiqCommAllGather(MugiqComm *comm, const void *send_d, void *recv_d, size_t bytes);
// Stream-ordered barrier: a one-element all-reduce enqueued on `stream` (cudaStream_t as void*); work enqueued on that
// stream afterwards starts only when every rank's earlier work on ITS stream has finished.  No host synchronisation.
void mugiqCommStreamBarrier(MugiqComm *comm, void *stream);
// eigenvector shard [lo, hi) of rank r: contiguous blocks whose sizes differ by at most one
void mugiqCommShard(int nEv, int rank, int size, int *lo, int *hi);

// Lattice-T split: Loop_Mugiq treats the eigenvectors as this rank's time slab of a lattice with comm_dim(3) slabs
// (communicator = getLoopComm(), all of whose ranks are time ranks; one rank = periodic in its own slab).
void setLoopTSplit(bool on);
bool getLoopTSplit();

// The communicator Loop_Mugiq sums its loop buffer over (nullptr = single process).
void setLoopComm(MugiqComm *comm);
MugiqComm *getLoopComm();

#endif
