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
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
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

int mugiq_b200_comm_destroy(mugiq_b200_comm_t *comm) {
  if (!comm) return MUGIQ_B200_OK;
  if (comm->side) {
    cudaStreamSynchronize(comm->side);
    cudaStreamDestroy(comm->side);
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

}  // namespace mugiq_b200

extern "C" {

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
  nchunks = std::max(1, std::min(nchunks, te - tb));
  std::vector<int> slots;
  for (const LoopPlan::Comp &c : pl.comps) slots.push_back(c.iL);
  for (int z : pl.zero_slots) (void)z;  // zero on every rank: nothing to sum
  int rc = MUGIQ_B200_OK;
  cudaEvent_t ev = nullptr;
  for (int k = 0; k < nchunks && rc == MUGIQ_B200_OK; k++) {
    const int t0 = tb + (int)((long long)(te - tb) * k / nchunks), t1 = tb + (int)((long long)(te - tb) * (k + 1) / nchunks);
    if ((rc = plan_accumulate_range(pl, dataPos_d, evec_d, sigma_h, nvec, accumulate, t0, t1, k == 0, stream))) break;
    if (comm->size == 1) continue;
    // chunk k is final on `stream`: its sum may start while the kernels of chunk k+1 run
    MUGIQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MUGIQ_CUDA_CHECK(cudaEventRecord(ev, stream));
    MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(comm->side, ev, 0));
    MUGIQ_CUDA_CHECK(cudaEventDestroy(ev));  // released when the recorded work has completed
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
