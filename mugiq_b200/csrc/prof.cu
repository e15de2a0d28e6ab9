// prof.cu — storage and C-ABI of the launch counters / event timers declared in prof.cuh.
#include <mutex>
#include <vector>

#include "fused.cuh"
#include "kernels.cuh"
#include "prof.cuh"

namespace mugiq_b200 {

namespace {
struct EventPair {
  cudaEvent_t a, b;
  int id;
};
struct ProfState {
  std::mutex mu;
  bool enabled = false;
  long long launches[K_COUNT] = {0};
  double bytes[K_COUNT] = {0};
  double flops[K_COUNT] = {0};
  double ms[K_COUNT] = {0};          // resolved time
  long long timed[K_COUNT] = {0};    // launches with resolved time
  std::vector<EventPair> pending;    // recorded, not yet resolved
  std::vector<EventPair> pool;       // free event pairs
  EventPair open[K_COUNT];
  bool is_open[K_COUNT] = {false};
};
ProfState &st() {
  static ProfState s;
  return s;
}
const char *kNames[K_COUNT] = {"contract_batch", "displace",      "loop_fused", "wilson_line", "minus_from_plus",
                               "reorder_mapgamma", "phase_matrix", "momproj",    "splitk_reduce", "convert_spinor", "halo_push",
                               "allreduce"};
}  // namespace

const char *kernel_name(int id) { return (id >= 0 && id < K_COUNT) ? kNames[id] : "?"; }

void prof_begin(int id, cudaStream_t stream, double alg_bytes, double alg_flops) {
  ProfState &s = st();
  std::lock_guard<std::mutex> lk(s.mu);
  s.launches[id]++;
  s.bytes[id] += alg_bytes;
  s.flops[id] += alg_flops;
  if (!s.enabled) return;
  EventPair ep;
  if (!s.pool.empty()) {
    ep = s.pool.back();
    s.pool.pop_back();
  } else {
    if (cudaEventCreate(&ep.a) != cudaSuccess || cudaEventCreate(&ep.b) != cudaSuccess) return;
  }
  ep.id = id;
  cudaEventRecord(ep.a, stream);
  s.open[id] = ep;
  s.is_open[id] = true;
}

void prof_end(int id, cudaStream_t stream) {
  ProfState &s = st();
  std::lock_guard<std::mutex> lk(s.mu);
  if (!s.is_open[id]) return;
  cudaEventRecord(s.open[id].b, stream);
  s.pending.push_back(s.open[id]);
  s.is_open[id] = false;
}

static void resolve_locked(ProfState &s) {
  for (EventPair &ep : s.pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(ep.b) == cudaSuccess && cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) {
      s.ms[ep.id] += ms;
      s.timed[ep.id]++;
    }
    s.pool.push_back(ep);
  }
  s.pending.clear();
}

}  // namespace mugiq_b200

using namespace mugiq_b200;

extern "C" {

int mugiq_b200_prof_enable(int on) {
  ProfState &s = st();
  std::lock_guard<std::mutex> lk(s.mu);
  s.enabled = on != 0;
  return MUGIQ_B200_OK;
}

int mugiq_b200_prof_reset(void) {
  ProfState &s = st();
  std::lock_guard<std::mutex> lk(s.mu);
  resolve_locked(s);
  for (int i = 0; i < K_COUNT; i++) {
    s.launches[i] = s.timed[i] = 0;
    s.bytes[i] = s.flops[i] = s.ms[i] = 0;
  }
  return MUGIQ_B200_OK;
}

int mugiq_b200_prof_num_kernels(void) { return K_COUNT; }

const char *mugiq_b200_prof_name(int kernel_id) { return kernel_name(kernel_id); }

int mugiq_b200_prof_query(int kernel_id, long long *launches, long long *timed_launches, double *ms_total,
                          double *alg_bytes_total, double *alg_flops_total) {
  if (kernel_id < 0 || kernel_id >= K_COUNT)
    return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_prof_query: kernel id %d out of range", kernel_id);
  ProfState &s = st();
  std::lock_guard<std::mutex> lk(s.mu);
  resolve_locked(s);
  if (launches) *launches = s.launches[kernel_id];
  if (timed_launches) *timed_launches = s.timed[kernel_id];
  if (ms_total) *ms_total = s.ms[kernel_id];
  if (alg_bytes_total) *alg_bytes_total = s.bytes[kernel_id];
  if (alg_flops_total) *alg_flops_total = s.flops[kernel_id];
  return MUGIQ_B200_OK;
}

int mugiq_b200_prof_fused_trace(void *trace_d, long long capacity_ctas) {
  if (trace_d && capacity_ctas < 1) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_prof_fused_trace: capacity %lld", capacity_ctas);
  fused_set_trace(static_cast<long long *>(trace_d), capacity_ctas);
  return MUGIQ_B200_OK;
}

}  // extern "C"
