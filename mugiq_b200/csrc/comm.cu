// comm.cu — the cross-GPU sum of the loop buffer (include/mugiq_b200.h, "eigenvector shards"): one process per GPU,
// every rank runs the loop plan on its shard of the eigenvectors, then the position-space buffer (or the projected one)
// is summed over the ranks with NCCL over NVLink 5 / NVSwitch.
//
// Replaces the reference's host-staged MPI_Reduce over COMM_SPACE / MPI_Gather over COMM_TIME / MPI_Bcast
// (/root/reference/lib/loop_mugiq.cpp:386-424: D2H copy, three host collectives) and its two MPI_Comm_split
// communicators (:62-88).
//
// NCCL is bound at run time (dlopen), not at link time: inside a PyTorch process the library must share torch's own NCCL
// instead of dragging a second copy in, and a process that never communicates needs no NCCL at all.
//
// mugiq_b200_loop_plan_accumulate_allreduce is the sharded step as ONE call: the plan's kernels run time-slice chunk by
// time-slice chunk on the caller's stream, and as soon as a chunk's loop values are final its all-reduce (one grouped
// NCCL launch over the chunk's contiguous runs: per loop, gamma and parity the sites of a time-slice range are
// V3/2 * nslices consecutive complex numbers) is issued on a high-priority side stream, so that the collective of chunk
// k travels over NVLink while the FP64-bound kernels of chunk k+1 compute; only the last chunk's sum is exposed.
// With peers attached (mugiq_b200_comm_attach_peers) the chunks travel by copy-engine pushes over IPC-mapped buffers instead
// (allreduce_pos_range_peer below): NCCL's all-reduce kernels and the loop kernels want the same SMs, the copy engines do not.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include "fused.cuh"
#include "plan.cuh"

namespace mugiq_b200 {

namespace {
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  bool ok = false;
};

// the NCCL already in the process (torch's) if there is one, else the system library
const NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
      if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
    for (const char *n : names)
      if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) return;
#define MUGIQ_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.handle, "nccl" #name))
    MUGIQ_NCCL_SYM(GetVersion);
    MUGIQ_NCCL_SYM(GetUniqueId);
    MUGIQ_NCCL_SYM(CommInitRank);
    MUGIQ_NCCL_SYM(CommDestroy);
    MUGIQ_NCCL_SYM(GetErrorString);
    MUGIQ_NCCL_SYM(AllReduce);
    MUGIQ_NCCL_SYM(AllGather);
    MUGIQ_NCCL_SYM(GroupStart);
    MUGIQ_NCCL_SYM(GroupEnd);
#undef MUGIQ_NCCL_SYM
    api.ok = api.GetVersion && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GetErrorString && api.AllReduce &&
             api.AllGather && api.GroupStart && api.GroupEnd;
  });
  return api.ok ? &api : nullptr;
}

int need_nccl(const NcclApi **api, const char *who) {
  *api = nccl_api();
  if (!*api) return set_error(MUGIQ_B200_ESTATE, "%s: no usable NCCL library (libnccl.so.2) in this process", who);
  return MUGIQ_B200_OK;
}
}  // namespace

#define MUGIQ_NCCL_CHECK(api, expr)                                                                                   \
  do {                                                                                                                \
    ncclResult_t r_ = (expr);                                                                                         \
    if (r_ != ncclSuccess)                                                                                            \
      return set_error(MUGIQ_B200_ECUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(r_)); \
  } while (0)

}  // namespace mugiq_b200

using namespace mugiq_b200;

struct mugiq_b200_comm_s {
  int rank = 0, size = 1, device = 0;
  ncclComm_t nccl = nullptr;
  cudaStream_t side = nullptr;  // high-priority stream of the overlapped all-reduces
  // peer transport (mugiq_b200_comm_attach_peers): every rank's position-space buffer and staging area mapped here
  std::vector<char *> peer_pos, peer_stage;
  size_t stage_bytes = 0;
  float *flag = nullptr;  // one element: the "my copies are done" all-reduce behind a round of copy-engine pushes
  unsigned epoch = 0;     // the staging area is used in halves, alternately
  static constexpr int kCopyStreams = 4;  // a round's copies fan out over these (several copy engines, latencies overlap)
  cudaStream_t copy[kCopyStreams] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[kCopyStreams] = {nullptr, nullptr, nullptr, nullptr};
};

extern "C" {

int mugiq_b200_comm_unique_id(void *id128) {
  const char *who = "mugiq_b200_comm_unique_id";
  if (!id128) return set_error(MUGIQ_B200_EINVAL, "%s: id128 is NULL", who);
  const NcclApi *api;
  int rc = need_nccl(&api, who);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == MUGIQ_B200_COMM_ID_BYTES, "ncclUniqueId size");
  MUGIQ_NCCL_CHECK(api, api->GetUniqueId(static_cast<ncclUniqueId *>(id128)));
  return MUGIQ_B200_OK;
}

int mugiq_b200_comm_create(mugiq_b200_comm_t **comm, const void *id128, int rank, int size) {
  const char *who = "mugiq_b200_comm_create";
  if (!comm) return set_error(MUGIQ_B200_EINVAL, "%s: comm is NULL", who);
  *comm = nullptr;
  if (size < 1 || rank < 0 || rank >= size) return set_error(MUGIQ_B200_EINVAL, "%s: bad rank/size %d/%d", who, rank, size);
  if (!id128) return set_error(MUGIQ_B200_EINVAL, "%s: id128 is NULL", who);
  const NcclApi *api;
  int rc = need_nccl(&api, who);
  if (rc) return rc;
  mugiq_b200_comm_s *c = new mugiq_b200_comm_s;
  c->rank = rank;
  c->size = size;
  if (cudaGetDevice(&c->device) != cudaSuccess) {
    delete c;
    return set_error(MUGIQ_B200_ECUDA, "%s: no CUDA device", who);
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclResult_t r = api->CommInitRank(&c->nccl, size, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return set_error(MUGIQ_B200_ECUDA, "%s: ncclCommInitRank failed: %s", who, api->GetErrorString(r));
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
  if (cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi) != cudaSuccess) {
    api->CommDestroy(c->nccl);
    delete c;
    return set_error(MUGIQ_B200_ECUDA, "%s: cannot create the communication stream", who);
  }
  *comm = c;
  return MUGIQ_B200_OK;
}

int mugiq_b200_comm_attach_peers(mugiq_b200_comm_t *comm, void *const *peer_pos_d, void *const *peer_stage_d, long long stage_bytes) {
  const char *who = "mugiq_b200_comm_attach_peers";
  if (!comm) return set_error(MUGIQ_B200_EINVAL, "%s: comm is NULL", who);
  if (comm->side) MUGIQ_CUDA_CHECK(cudaStreamSynchronize(comm->side));
  comm->peer_pos.clear();
  comm->peer_stage.clear();
  comm->stage_bytes = 0;
  if (!peer_pos_d && !peer_stage_d) return MUGIQ_B200_OK;  // detach
  if (!peer_pos_d || !peer_stage_d || stage_bytes < 1) return set_error(MUGIQ_B200_EINVAL, "%s: NULL table or empty staging area", who);
  for (int r = 0; r < comm->size; r++) {
    if (!peer_pos_d[r] || !peer_stage_d[r]) return set_error(MUGIQ_B200_EINVAL, "%s: rank %d has a NULL buffer", who, r);
    if (((uintptr_t)peer_pos_d[r] | (uintptr_t)peer_stage_d[r]) & 255)
      return set_error(MUGIQ_B200_EINVAL, "%s: the buffers of rank %d are not 256-byte aligned", who, r);
  }
  if (!comm->flag) {
    MUGIQ_CUDA_CHECK(cudaMalloc((void **)&comm->flag, sizeof(float)));
    MUGIQ_CUDA_CHECK(cudaMemset(comm->flag, 0, sizeof(float)));
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    MUGIQ_CUDA_CHECK(cudaEventCreateWithFlags(&comm->fork, cudaEventDisableTiming));
    for (int k = 0; k < mugiq_b200_comm_s::kCopyStreams; k++) {
      MUGIQ_CUDA_CHECK(cudaStreamCreateWithPriority(&comm->copy[k], cudaStreamNonBlocking, hi));
      MUGIQ_CUDA_CHECK(cudaEventCreateWithFlags(&comm->join[k], cudaEventDisableTiming));
    }
  }
  for (int r = 0; r < comm->size; r++) {
    comm->peer_pos.push_back(static_cast<char *>(peer_pos_d[r]));
    comm->peer_stage.push_back(static_cast<char *>(peer_stage_d[r]));
  }
  comm->stage_bytes = (size_t)stage_bytes;
  return MUGIQ_B200_OK;
}

int mugiq_b200_comm_destroy(mugiq_b200_comm_t *comm) {
  if (!comm) return MUGIQ_B200_OK;
  if (comm->side) {
    cudaStreamSynchronize(comm->side);
    cudaStreamDestroy(comm->side);
  }
  if (comm->flag) cudaFree(comm->flag);
  if (comm->fork) cudaEventDestroy(comm->fork);
  for (int k = 0; k < mugiq_b200_comm_s::kCopyStreams; k++) {
    if (comm->copy[k]) {
      cudaStreamSynchronize(comm->copy[k]);
      cudaStreamDestroy(comm->copy[k]);
    }
    if (comm->join[k]) cudaEventDestroy(comm->join[k]);
  }
  if (comm->nccl)
    if (const NcclApi *api = nccl_api()) api->CommDestroy(comm->nccl);
  delete comm;
  return MUGIQ_B200_OK;
}

int mugiq_b200_comm_info(const mugiq_b200_comm_t *comm, int *rank, int *size, int *nccl_version) {
  if (!comm) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_comm_info: comm is NULL");
  if (rank) *rank = comm->rank;
  if (size) *size = comm->size;
  if (nccl_version) {
    *nccl_version = 0;
    if (const NcclApi *api = nccl_api()) api->GetVersion(nccl_version);
  }
  return MUGIQ_B200_OK;
}

int mugiq_b200_allreduce(void *buf_d, long long count, int precision, mugiq_b200_comm_t *comm, void *stream) {
  const char *who = "mugiq_b200_allreduce";
  if (!comm) return set_error(MUGIQ_B200_EINVAL, "%s: comm is NULL", who);
  if (!buf_d) return set_error(MUGIQ_B200_EINVAL, "%s: buf_d is NULL", who);
  if (count < 0) return set_error(MUGIQ_B200_EINVAL, "%s: count = %lld", who, count);
  if (precision != MUGIQ_B200_PREC_SINGLE && precision != MUGIQ_B200_PREC_DOUBLE)
    return set_error(MUGIQ_B200_EINVAL, "%s: precision %d not supported", who, precision);
  if (count == 0) return MUGIQ_B200_OK;
  const NcclApi *api;
  int rc = need_nccl(&api, who);
  if (rc) return rc;
  ProfScope prof(K_ALLREDUCE, (cudaStream_t)stream, (double)count * prec_bytes(precision));
  MUGIQ_NCCL_CHECK(api, api->AllReduce(buf_d, buf_d, (size_t)count, precision == MUGIQ_B200_PREC_DOUBLE ? ncclDouble : ncclFloat,
                                       ncclSum, comm->nccl, (cudaStream_t)stream));
  return MUGIQ_B200_OK;
}

int mugiq_b200_allgather(void *recv_d, const void *send_d, long long bytes, mugiq_b200_comm_t *comm, void *stream) {
  const char *who = "mugiq_b200_allgather";
  if (!comm) return set_error(MUGIQ_B200_EINVAL, "%s: comm is NULL", who);
  if (!recv_d || !send_d) return set_error(MUGIQ_B200_EINVAL, "%s: NULL buffer", who);
  if (bytes < 0) return set_error(MUGIQ_B200_EINVAL, "%s: bytes = %lld", who, bytes);
  if (bytes == 0) return MUGIQ_B200_OK;
  const NcclApi *api;
  int rc = need_nccl(&api, who);
  if (rc) return rc;
  MUGIQ_NCCL_CHECK(api, api->AllGather(send_d, recv_d, (size_t)bytes, ncclChar, comm->nccl, (cudaStream_t)stream));
  return MUGIQ_B200_OK;
}

}  // extern "C"

namespace mugiq_b200 {

// all-reduce of the time-slices [t0, t1) of the given loop slots: 16 gammas x 2 parities contiguous runs per slot, one
// grouped NCCL launch
static int allreduce_pos_range(void *dataPos_d, const int *slots, int nslots, int t0, int t1, const LatGeom &g, int precision,
                               mugiq_b200_comm_t *comm, cudaStream_t stream, const char *who) {
  const NcclApi *api;
  int rc = need_nccl(&api, who);
  if (rc) return rc;
  const size_t pb = prec_bytes(precision);
  const size_t V3h = (size_t)g.V3 / 2;
  const size_t run = V3h * (size_t)(t1 - t0) * 2;  // real numbers per piece
  const ncclDataType_t dt = precision == MUGIQ_B200_PREC_DOUBLE ? ncclDouble : ncclFloat;
  char *base = static_cast<char *>(dataPos_d);
  const bool whole = t0 == 0 && t1 == g.L[3];
  ProfScope prof(K_ALLREDUCE, stream, (double)nslots * 16 * 2 * run * pb);
  MUGIQ_NCCL_CHECK(api, api->GroupStart());
  ncclResult_t r = ncclSuccess;
  for (int s = 0; s < nslots && r == ncclSuccess; s++) {
    if (whole) {  // the slot is one contiguous block
      char *p = base + (size_t)slots[s] * 16 * g.volume * 2 * pb;
      r = api->AllReduce(p, p, (size_t)16 * g.volume * 2, dt, ncclSum, comm->nccl, stream);
      continue;
    }
    for (int G = 0; G < 16 && r == ncclSuccess; G++)
      for (int par = 0; par < 2 && r == ncclSuccess; par++) {
        char *p = base + (((size_t)slots[s] * 16 + G) * g.volume + (size_t)par * g.volumeCB + (size_t)t0 * V3h) * 2 * pb;
        r = api->AllReduce(p, p, run, dt, ncclSum, comm->nccl, stream);
      }
  }
  ncclResult_t e = api->GroupEnd();
  if (r != ncclSuccess) return set_error(MUGIQ_B200_ECUDA, "%s: ncclAllReduce failed: %s", who, api->GetErrorString(r));
  if (e != ncclSuccess) return set_error(MUGIQ_B200_ECUDA, "%s: ncclGroupEnd failed: %s", who, api->GetErrorString(e));
  return MUGIQ_B200_OK;
}

// Time-slice chunks of the overlapped sum: each chunk is half of what is left, down to (te - tb) / (2 nchunks) slices.  The
// sum of chunk k hides under the kernels of chunk k+1 as long as those take longer (moving a slice costs ~1/6 of computing
// it at configs[3] on 8 GPUs), only the LAST chunk's sum is exposed - so the last chunk is small - and every chunk is a
// separate set of kernel launches with its own tail - so there are few (64 slices, nchunks = 8: 32, 16, 8, 4, 4).
static std::vector<int> chunk_bounds(int tb, int te, int nchunks) {
  std::vector<int> b{tb};
  const int T = te - tb;
  if (nchunks <= 1 || T <= 1) {
    b.push_back(te);
    return b;
  }
  const int minsz = std::max(1, T / (2 * nchunks));
  int left = T;
  while (left > 0) {
    int c = std::max(minsz, left / 2);
    if (left - c < minsz) c = left;
    b.push_back(b.back() + c);
    left -= c;
  }
  return b;
}

// ---- the same sum over peer memory, moved by the copy engines ------------------------------------------------------------
// NCCL's all-reduce kernels need SMs, and a fused CTA owns all registers and shared memory of its SM: chunk k's collective
// only progresses where chunk k+1's kernels leave SMs free, and what it takes it takes from them (8 GPUs, configs[3]: 3.7 ms
// of a 36.5 ms step stay exposed out of 14.9 ms of NCCL kernel time).  Here the bytes travel by DMA: every rank maps all
// position-space buffers and a staging area per rank (CUDA IPC).  The chunk's pieces - one per (slot, gamma, parity): the
// sites of the time-slice range are one contiguous run - are dealt round-robin to owner ranks;
//   1. reduce-scatter: 2-D copies push my copy of every piece into its owner's staging area (strided in, dense out);
//   2. a one-element NCCL all-reduce behind the pushes tells every rank that all contributions have landed;
//   3. a small kernel adds the size-1 staged copies of the pieces I own to mine;
//   4. all-gather: 2-D copies push the summed pieces straight into every peer's position-space buffer.
// Only step 3 and the one-element all-reduce run on SMs.  After the last chunk a second one-element all-reduce closes the
// all-gather.  The staging area is used in halves: chunk k+1's pushes cannot land before every rank has finished step 2 of
// chunk k+1, i.e. issued step 3 of chunk k.
constexpr int kPeerMaxSlots = 256;  // computed loops of a plan (the table travels as a kernel parameter)
struct PeerSumArgs {
  const char *stage;  // this rank's staging area (half in use)
  int slots[kPeerMaxSlots];
  int nslots, W, me, M;
  long long piece_elems;  // reals between two pieces of a slot in dataPos (volumeCB * 2)
  long long t0_elems;     // reals before the time-slice range inside a piece
  long long run_elems;    // reals per piece of this chunk
  long long run_cap;      // bytes per piece in the staging area
};
template <typename F> __global__ void __launch_bounds__(256) peer_sum_kernel(F *__restrict__ pos, const PeerSumArgs a) {
  const int m = blockIdx.y;
  const long long q = a.me + (long long)a.W * m;
  if (q >= 32LL * a.nslots) return;
  const int s = (int)(q / 32), i = (int)(q % 32);
  F *dst = pos + ((long long)a.slots[s] * 32 + i) * a.piece_elems + a.t0_elems;
  using V = typename vec2_of<F>::type;
  const long long n2 = a.run_elems / 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += (long long)gridDim.x * blockDim.x) {
    V acc = reinterpret_cast<const V *>(dst)[e];
    for (int j = 0; j < a.W; j++) {
      if (j == a.me) continue;
      const V v = reinterpret_cast<const V *>(a.stage + ((size_t)j * a.M + m) * a.run_cap)[e];
      acc.x += v.x;
      acc.y += v.y;
    }
    reinterpret_cast<V *>(dst)[e] = acc;
  }
}

static int allreduce_pos_range_peer(void *dataPos_d, const int *slots, int nslots, int t0, int t1, const LatGeom &g, int precision,
                                    mugiq_b200_comm_t *comm, cudaStream_t side, bool last, const char *who) {
  const NcclApi *api;
  int rc = need_nccl(&api, who);
  if (rc) return rc;
  const int W = comm->size, me = comm->rank;
  const size_t pb = prec_bytes(precision);
  const size_t V3h = (size_t)g.V3 / 2;
  const size_t run_bytes = V3h * (size_t)(t1 - t0) * 2 * pb;
  const size_t piece_bytes = (size_t)g.volumeCB * 2 * pb;  // between the (gamma, parity) pieces of a slot
  const size_t t0_bytes = (size_t)t0 * V3h * 2 * pb;
  const int P = nslots * 32, M = (P + W - 1) / W;
  const size_t run_cap = comm->stage_bytes / ((size_t)2 * W * M) / 256 * 256;
  if (run_bytes > run_cap)
    return set_error(MUGIQ_B200_EINVAL, "%s: the staging area holds %zu bytes per piece, a chunk of %d time-slices needs %zu (see mugiq_b200_comm_stage_bytes)",
                     who, run_cap, t1 - t0, run_bytes);
  if (nslots > kPeerMaxSlots) return set_error(MUGIQ_B200_EINVAL, "%s: %d computed loops, the peer transport takes %d", who, nslots, kPeerMaxSlots);
  const size_t half = (size_t)(comm->epoch++ & 1) * W * M * run_cap;
  char *pos = static_cast<char *>(dataPos_d);
  ProfScope prof(K_ALLREDUCE, side, (double)nslots * 32 * run_bytes);
  // rows of slot s (index s_idx in the list) that belong to owner o: i = i0, i0 + W, ... < 32
  auto rows = [&](int s_idx, int o, int *i0, int *count) {
    *i0 = (((o - s_idx * 32) % W) + W) % W;
    *count = *i0 < 32 ? (32 - *i0 + W - 1) / W : 0;
  };
  // a round of copies: what `side` has done so far precedes them, `side` continues when they are all done; peer d's copies
  // go to copy stream d mod kCopyStreams
  constexpr int KS = mugiq_b200_comm_s::kCopyStreams;
  auto fork = [&]() -> int {
    MUGIQ_CUDA_CHECK(cudaEventRecord(comm->fork, side));
    for (int k = 0; k < KS; k++) MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(comm->copy[k], comm->fork, 0));
    return MUGIQ_B200_OK;
  };
  auto join = [&]() -> int {
    for (int k = 0; k < KS; k++) {
      MUGIQ_CUDA_CHECK(cudaEventRecord(comm->join[k], comm->copy[k]));
      MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(side, comm->join[k], 0));
    }
    return MUGIQ_B200_OK;
  };
  if ((rc = fork())) return rc;
  for (int d = 1; d < W; d++) {  // 1. my pieces to their owners (start with my right-hand neighbour: the ranks fan out)
    const int o = (me + d) % W;
    for (int s = 0; s < nslots; s++) {
      int i0, count;
      rows(s, o, &i0, &count);
      if (!count) continue;
      const size_t m0 = ((size_t)s * 32 + i0) / W;
      MUGIQ_CUDA_CHECK(cudaMemcpy2DAsync(comm->peer_stage[o] + half + ((size_t)me * M + m0) * run_cap, run_cap,
                                         pos + ((size_t)slots[s] * 32 + i0) * piece_bytes + t0_bytes, (size_t)W * piece_bytes, run_bytes,
                                         count, cudaMemcpyDeviceToDevice, comm->copy[d % KS]));
    }
  }
  if ((rc = join())) return rc;
  const ncclDataType_t ft = ncclFloat;
  MUGIQ_NCCL_CHECK(api, api->AllReduce(comm->flag, comm->flag, 1, ft, ncclSum, comm->nccl, side));  // 2.
  {  // 3.
    PeerSumArgs a;
    a.stage = comm->peer_stage[me] + half;
    for (int s = 0; s < nslots; s++) a.slots[s] = slots[s];
    a.nslots = nslots;
    a.W = W;
    a.me = me;
    a.M = M;
    a.piece_elems = (long long)g.volumeCB * 2;
    a.t0_elems = (long long)t0 * (long long)V3h * 2;
    a.run_elems = (long long)V3h * (t1 - t0) * 2;
    a.run_cap = (long long)run_cap;
    const dim3 grid((unsigned)std::min<long long>((a.run_elems / 2 + 255) / 256, 64), (unsigned)M);
    if (precision == MUGIQ_B200_PREC_DOUBLE)
      peer_sum_kernel<double><<<grid, 256, 0, side>>>(static_cast<double *>(dataPos_d), a);
    else
      peer_sum_kernel<float><<<grid, 256, 0, side>>>(static_cast<float *>(dataPos_d), a);
    MUGIQ_LAUNCH_CHECK();
  }
  if ((rc = fork())) return rc;
  for (int d = 1; d < W; d++) {  // 4. the pieces I own, summed, into every peer's buffer
    const int p = (me + d) % W;
    for (int s = 0; s < nslots; s++) {
      int i0, count;
      rows(s, me, &i0, &count);
      if (!count) continue;
      const size_t off = ((size_t)slots[s] * 32 + i0) * piece_bytes + t0_bytes;
      MUGIQ_CUDA_CHECK(cudaMemcpy2DAsync(comm->peer_pos[p] + off, (size_t)W * piece_bytes, pos + off, (size_t)W * piece_bytes, run_bytes,
                                         count, cudaMemcpyDeviceToDevice, comm->copy[d % KS]));
    }
  }
  if ((rc = join())) return rc;
  if (last) MUGIQ_NCCL_CHECK(api, api->AllReduce(comm->flag, comm->flag, 1, ft, ncclSum, comm->nccl, side));
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200

extern "C" {

long long mugiq_b200_comm_stage_bytes(const mugiq_b200_loop_plan_t *plan, int nchunks, int size) {
  if (!plan || size < 1) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_comm_stage_bytes: bad argument");
  const LoopPlan &pl = plan_of(plan);
  const int tb = pl.t_begin, te = pl.t_end < 0 ? pl.g.L[3] : pl.t_end;
  const std::vector<int> cb = chunk_bounds(tb, te, nchunks);
  long long slices = 1;  // the longest chunk
  for (size_t k = 0; k + 1 < cb.size(); k++) slices = std::max<long long>(slices, cb[k + 1] - cb[k]);
  const long long run = ((long long)pl.g.V3 / 2 * slices * 2 * (long long)prec_bytes(pl.precision) + 255) / 256 * 256;
  const long long P = 32LL * (long long)pl.comps.size(), M = (P + size - 1) / size;
  return 2 * (long long)size * M * run;
}

int mugiq_b200_allreduce_pos(void *dataPos_d, const int *slots_h, int nslots, int t_begin, int t_end,
                             const mugiq_b200_geom_t *geom, mugiq_b200_comm_t *comm, void *stream) {
  const char *who = "mugiq_b200_allreduce_pos";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if (!comm) return set_error(MUGIQ_B200_EINVAL, "%s: comm is NULL", who);
  if (!dataPos_d || !slots_h) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  if (nslots < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nslots = %d", who, nslots);
  if (t_end < 0) t_end = geom->L[3];
  if (t_begin < 0 || t_begin >= t_end || t_end > geom->L[3])
    return set_error(MUGIQ_B200_EINVAL, "%s: bad time-slice range [%d, %d) on Lt = %d", who, t_begin, t_end, geom->L[3]);
  for (int s = 0; s < nslots; s++)
    if (slots_h[s] < 0) return set_error(MUGIQ_B200_EINVAL, "%s: slot %d is negative", who, s);
  return allreduce_pos_range(dataPos_d, slots_h, nslots, t_begin, t_end, make_geom(geom->L), geom->precision, comm,
                             (cudaStream_t)stream, who);
}

int mugiq_b200_loop_plan_accumulate_allreduce(const mugiq_b200_loop_plan_t *plan, void *dataPos_d, const void *const *evec_d,
                                              const double *sigma_h, int nvec, int accumulate, mugiq_b200_comm_t *comm,
                                              int nchunks, void *stream_) {
  const char *who = "mugiq_b200_loop_plan_accumulate_allreduce";
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "%s: plan is NULL", who);
  if (!comm) return set_error(MUGIQ_B200_EINVAL, "%s: comm is NULL", who);
  if (!dataPos_d || !evec_d || !sigma_h) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  if (nvec < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nvec = %d must be positive", who, nvec);
  for (int i = 0; i < nvec; i++)
    if (!evec_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: eigenvector %d is NULL", who, i);
  const LoopPlan &pl = plan_of(plan);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int tb = pl.t_begin, te = pl.t_end < 0 ? pl.g.L[3] : pl.t_end;
  const std::vector<int> cb = chunk_bounds(tb, te, nchunks);
  nchunks = (int)cb.size() - 1;
  std::vector<int> slots;
  for (const LoopPlan::Comp &c : pl.comps) slots.push_back(c.iL);
  for (int z : pl.zero_slots) (void)z;  // zero on every rank: nothing to sum
  int rc = MUGIQ_B200_OK;
  cudaEvent_t ev = nullptr;
  for (int k = 0; k < nchunks && rc == MUGIQ_B200_OK; k++) {
    const int t0 = cb[k], t1 = cb[k + 1];
    if ((rc = plan_accumulate_range(pl, dataPos_d, evec_d, sigma_h, nvec, accumulate, t0, t1, k == 0, stream))) break;
    if (comm->size == 1) continue;
    // chunk k is final on `stream`: its sum may start while the kernels of chunk k+1 run
    MUGIQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MUGIQ_CUDA_CHECK(cudaEventRecord(ev, stream));
    MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(comm->side, ev, 0));
    MUGIQ_CUDA_CHECK(cudaEventDestroy(ev));  // released when the recorded work has completed
    // peers attached and this is the buffer they map: the bytes travel by DMA, else through NCCL's kernels
    if (!comm->peer_pos.empty() && comm->peer_pos[comm->rank] == static_cast<char *>(dataPos_d))
      rc = allreduce_pos_range_peer(dataPos_d, slots.data(), (int)slots.size(), t0, t1, pl.g, pl.precision, comm, comm->side,
                                    k == nchunks - 1, who);
    else
      rc = allreduce_pos_range(dataPos_d, slots.data(), (int)slots.size(), t0, t1, pl.g, pl.precision, comm, comm->side, who);
  }
  if (rc == MUGIQ_B200_OK && comm->size > 1) {  // what follows on `stream` sees the summed buffer
    MUGIQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MUGIQ_CUDA_CHECK(cudaEventRecord(ev, comm->side));
    MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(stream, ev, 0));
    MUGIQ_CUDA_CHECK(cudaEventDestroy(ev));
  }
  return rc;
}

}  // extern "C"
